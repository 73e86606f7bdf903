"""GPU, BASELINE.json sizes (480x640 input: stage-0 maps 120x160, d_inner 192, d_state 16 / 4): the CPU oracle does
not finish these in seconds, so parity is checked through size-independent properties of the path:
  * direction symmetry: with one weight set for all 4 directions, the reversed / column-major scans of an image
    are the row-major scan of the flipped / transposed image (the definition of CrossScan, vmamba.py:80-97);
  * linearity in u at fixed delta, B, C (selective_scan_interface.py:100-131: y = sum_n C·h + D·u, h linear in u);
  * L-segment invariance: MODE_SUMMARY -> combine -> MODE_APPLY equals the single serial pass;
  * two independent implementations agree: the fused channels-last path (TMA direction walks + tcgen05 GEMMs) against
    the composed path (CrossScan index shuffles + the op-level kernel that is pinned to the reference's goldens),
    for one SS2D block at stage-0 size and for the whole Sigma-tiny network at 480x640 (logits, arg-max labels)."""
import contextlib
import io

import numpy as np
import pytest
import torch

import procedural as P
from helpers import cfg_tiny

pytestmark = pytest.mark.gpu
S = 33
H, W, D, R = 120, 160, 192, 6


def _scan_inputs(N, Bt, same_weights=True):
    from sigma_b200 import _lib
    Cp = _lib.lib().sigma_ss2d_padded_cp(N, R)
    g = torch.Generator(device="cuda").manual_seed(S)
    L = H * W
    xc = torch.randn(Bt, L, D, device="cuda", generator=g)
    xdbl = torch.randn(Bt, L, 4, Cp, device="cuda", generator=g)
    xdbl[..., 2 * N + R:] = 0.0
    if same_weights:
        xdbl[:, :, 1:] = xdbl[:, :, :1]
    dtw = ((torch.rand(1, D, R, device="cuda", generator=g) * 2 - 1) * R ** -0.5).expand(4, D, R).contiguous()
    dtb = (torch.rand(1, D, device="cuda", generator=g) * 5 - 6).expand(4, D).contiguous()
    A = (-(torch.rand(1, D, N, device="cuda", generator=g) * N + 0.3)).expand(4, D, N).reshape(4 * D, N).contiguous()
    Ds = torch.randn(1, D, device="cuda", generator=g).expand(4, D).reshape(-1).contiguous()
    return Cp, xc, xdbl, dtw, dtb, A, Ds


def _scan(xc, xdbl, dtw, dtb, A, Ds, N, Cp, h=H, w=W, split=0):
    from sigma_b200 import _lib, fused
    fused._FORCE_SPLIT = split
    try:
        return fused.ss2d_scan(_lib.DIRS_CROSS4, xc.contiguous(), xdbl.contiguous(), dtw, dtb, A, Ds, xc.shape[0], h, w, D, N, R, Cp)
    finally:
        fused._FORCE_SPLIT = 0


@pytest.mark.parametrize("N", [16, 4])
def test_direction_symmetry_fullsize(N):
    """y[k] is stored at the position it belongs to, so with shared weights: y2 = flipL(y0 on flipL(x)), y1 = y0 on the
    transposed image (read back transposed), y3 = the same for the reversed column-major walk."""
    Cp, xc, xdbl, dtw, dtb, A, Ds = _scan_inputs(N, 2)
    y = _scan(xc, xdbl, dtw, dtb, A, Ds, N, Cp)                                   # (4, B, L, D)
    scale = float(y.abs().max())
    yf = _scan(xc.flip(1), xdbl.flip(1), dtw, dtb, A, Ds, N, Cp)                 # image reversed along L
    assert float((y[2] - yf[0].flip(1)).abs().max()) <= 2e-4 * scale, "reversed direction != forward scan of the reversed sequence"
    B = xc.shape[0]
    xt = xc.view(B, H, W, D).transpose(1, 2).reshape(B, H * W, D)                 # transposed image: (W, H) row-major
    dt = xdbl.view(B, H, W, 4, Cp).transpose(1, 2).reshape(B, H * W, 4, Cp)
    yt = _scan(xt, dt, dtw, dtb, A, Ds, N, Cp, h=W, w=H)
    back = lambda t: t.view(B, W, H, D).transpose(1, 2).reshape(B, H * W, D)
    assert float((y[1] - back(yt[0])).abs().max()) <= 2e-4 * scale, "column-major direction != row-major scan of the transposed image"
    assert float((y[3] - back(yt[2])).abs().max()) <= 2e-4 * scale, "reversed column-major direction != reversed scan of the transposed image"


@pytest.mark.parametrize("N", [16, 4])
def test_linearity_in_u_and_split_invariance_fullsize(N):
    Cp, xc, xdbl, dtw, dtb, A, Ds = _scan_inputs(N, 2, same_weights=False)
    x2 = torch.randn_like(xc)
    y1 = _scan(xc, xdbl, dtw, dtb, A, Ds, N, Cp)
    y2 = _scan(x2, xdbl, dtw, dtb, A, Ds, N, Cp)
    y12 = _scan(0.5 * xc - 2.0 * x2, xdbl, dtw, dtb, A, Ds, N, Cp)
    scale = float(max(y1.abs().max(), y2.abs().max()))
    assert float((y12 - (0.5 * y1 - 2.0 * y2)).abs().max()) <= 5e-5 * scale, "the scan is not linear in u"
    for split in (2, 5):
        ys = _scan(xc, xdbl, dtw, dtb, A, Ds, N, Cp, split=split)
        assert float((ys - y1).abs().max()) <= 2e-4 * scale, f"{split} L-segments differ from the single pass"


def test_ss2d_block_fused_vs_composed_stage0():
    """One VSSBlock at the stage-0 size of the 480x640 workload (B·L = 2·19200 rows, d_inner 192, d_state 16)."""
    from sigma_b200 import modules as M
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(S)
    blk = M.VSSBlock(hidden_dim=96, drop_path=0.0, d_state=16, ssm_ratio=2.0, dt_rank="auto", mlp_ratio=0.0).cuda().eval()
    P.fill_state_dict(blk, S)
    x = torch.randn(2, H, W, 96, device="cuda")
    with torch.no_grad():
        got = blk(x)
        with M.composed_path():
            ref = blk(x)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 1e-2 * scale, f"fused vs composed VSSBlock: {float((got - ref).abs().max()):.3e} of {scale:.3e}"


def test_sigma_tiny_480x640_fused_vs_composed():
    """The whole north-star forward at BASELINE size, one image: logits of the fused path against the composed path."""
    from sigma_b200 import modules as M
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(S)
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.EncoderDecoder(cfg_tiny(480, 640), criterion=None).cuda().eval()
    P.fill_state_dict(model, S)
    rgb = torch.randn(1, 3, 480, 640, device="cuda")
    mx = torch.randn(1, 3, 480, 640, device="cuda")
    with torch.no_grad():
        got = model(rgb, mx)
        with M.composed_path():
            ref = model(rgb, mx)
    assert got.shape == ref.shape == (1, 9, 480, 640)
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    agree = float((got.argmax(1) == ref.argmax(1)).float().mean())
    assert err <= 1e-2 * scale, f"logits differ by {err:.3e} (scale {scale:.3e})"
    assert agree >= 0.99, f"only {agree:.4f} of the arg-max labels agree"


@pytest.mark.parametrize("backbone,Hh,Ww", [("sigma_small", 96, 128), ("sigma_base", 96, 160)])
def test_other_backbones_fused_vs_composed(backbone, Hh, Ww):
    """BASELINE.json configs 3-5 name Sigma-small and Sigma-base (dims 128..1024, dt_rank 8..64, 27-deep stage 3): the fused
    path must agree with the composed path (op-level kernel pinned to the reference's goldens) for them as well."""
    from sigma_b200 import modules as M
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(S)
    with contextlib.redirect_stdout(io.StringIO()):
        model = M.EncoderDecoder(cfg_tiny(Hh, Ww, num_classes=40, backbone=backbone), criterion=None).cuda().eval()
    P.fill_state_dict(model, S)
    rgb = torch.randn(1, 3, Hh, Ww, device="cuda")
    mx = torch.randn(1, 3, Hh, Ww, device="cuda")
    with torch.no_grad():
        got = model(rgb, mx)
        with M.composed_path():
            ref = model(rgb, mx)
    assert got.shape == ref.shape == (1, 40, Hh, Ww)
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    agree = float((got.argmax(1) == ref.argmax(1)).float().mean())
    assert err <= 2e-2 * scale, f"{backbone}: logits differ by {err:.3e} (scale {scale:.3e})"
    assert agree >= 0.97, f"{backbone}: only {agree:.4f} of the arg-max labels agree"
