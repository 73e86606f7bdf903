"""GPU: parity against the reference CUDA EXTENSION itself (selective_scan_cuda_core rebuilt for sm_100a by
baseline/build_ref_ext.py into baseline/_ref/, BASELINE.json north_star: "outputs match the reference selective_scan
extension on identical inputs within 1e-3 fp32 / 1e-2 bf16").  The extension is built with --use_fast_math (setup.py:94);
errors are measured against the output scale.  Skipped when the built .so did not travel (it is git-ignored)."""
import os
import sys

import numpy as np
import pytest
import torch

import procedural as P
from test_ss2d_scan_gpu import _dir_index

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
S = 41


@pytest.fixture(scope="module")
def ext():
    if not os.path.exists(os.path.join(REF_DIR, "selective_scan_cuda_core.so")):
        pytest.skip("baseline/_ref/selective_scan_cuda_core.so not built (python baseline/build_ref_ext.py)")
    sys.path.insert(0, REF_DIR)
    import selective_scan_cuda_core as m
    return m


def _inputs(b, d, n, L, g, dt):
    tag = f"ext/{b}/{d}/{n}/{L}/{g}"
    u = P.randn(S, tag + "/u", (b, d, L)).to(dt).cuda()
    dl = P.randn(S, tag + "/dl", (b, d, L), 0.7).to(dt).cuda()
    A = (-P.rand(S, tag + "/A", (d, n), 0.3, n + 0.5)).cuda()
    Bm = P.randn(S, tag + "/B", (b, g, n, L)).to(dt).cuda()
    Cm = P.randn(S, tag + "/C", (b, g, n, L)).to(dt).cuda()
    D = P.randn(S, tag + "/D", (d,)).cuda()
    bias = P.rand(S, tag + "/bias", (d,), -6.0, -2.0).cuda()
    return u, dl, A, Bm, Cm, D, bias


SHAPES = [  # b, d, n, L, g: Sigma shapes (App. B) and ragged ones
    (2, 3072, 16, 1200, 4), (1, 768, 16, 4800, 4), (2, 6144, 16, 300, 4), (2, 768, 4, 1200, 1), (1, 384, 4, 9600, 2),
    (2, 3072, 4, 1200, 4), (2, 24, 8, 372, 2), (1, 40, 16, 2077, 1), (3, 36, 4, 4100, 3), (1, 8192, 16, 690, 4),
]


@pytest.mark.parametrize("b,d,n,L,g", SHAPES)
@pytest.mark.parametrize("dn", ["f32", "bf16", "f16"])
def test_fwd_matches_extension(ext, b, d, n, L, g, dn):
    from sigma_b200 import ops
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dn]
    u, dl, A, Bm, Cm, D, bias = _inputs(b, d, n, L, g, dt)
    ref, xr = ext.fwd(u, dl, A, Bm, Cm, D, bias, True, 1)
    got, xg = ops.selective_scan_cuda_core_fwd(u, dl, A, Bm, Cm, D, bias, True, 1)
    assert got.dtype == ref.dtype and got.shape == ref.shape and xg.shape == xr.shape
    scale = float(ref.float().abs().max())
    err = float((got.float() - ref.float()).abs().max()) / scale
    assert err <= (1e-3 if dn == "f32" else 1e-2), f"fwd vs extension ({dn}): {err:.3e} of the output scale"
    if dn == "f32":   # chunk-end states (only .y = h is consumed, selective_scan_bwd_kernel.cuh:114-116); running prefix in .x
        xs = float(xr[..., 1::2].abs().max()) + 1e-20
        assert float((xg[..., 1::2] - xr[..., 1::2]).abs().max()) / xs <= 1e-3
        assert float((xg[..., 0::2] - xr[..., 0::2]).abs().max()) <= 1e-3, "running-prefix component of x"


@pytest.mark.parametrize("b,d,n,L,g", SHAPES[:8])
@pytest.mark.parametrize("dn", ["f32", "bf16", "f16"])
def test_bwd_matches_extension(ext, b, d, n, L, g, dn):
    from sigma_b200 import ops
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dn]
    u, dl, A, Bm, Cm, D, bias = _inputs(b, d, n, L, g, dt)
    dout = P.randn(S, f"ext/dout/{b}/{d}/{L}", (b, d, L)).to(dt).cuda()
    _, xr = ext.fwd(u, dl, A, Bm, Cm, D, bias, True, 1)
    ref = ext.bwd(u, dl, A, Bm, Cm, D, bias, dout, xr, True, 1)
    got = ops.selective_scan_cuda_core_bwd(u, dl, A, Bm, Cm, D, bias, dout, None, True, 1)
    bar = 2e-3 if dn == "f32" else 2e-2   # fast-math reference, atomics in a different order
    for nm, r, o in zip(("du", "ddelta", "dA", "dB", "dC", "dD", "dbias"), ref, got):
        assert o.dtype == r.dtype and o.shape == r.shape, nm
        sc = float(r.float().abs().max()) + 1e-20
        err = float((o.float() - r.float()).abs().max()) / sc
        assert err <= bar, f"{nm} vs extension ({dn}): {err:.3e} of its scale"


@pytest.mark.parametrize("kind,B,H,W,D,N,R", [("cross4", 2, 30, 40, 768, 16, 24), ("cross4", 1, 60, 80, 384, 4, 12),
                                              ("seq2", 1, 30, 40, 768, 4, 24), ("cross4", 1, 23, 30, 256, 16, 16)])
def test_fused_scan_matches_extension(ext, kind, B, H, W, D, N, R):
    """The fused channels-last scan against CrossScan + dt einsum (fp32) + the EXTENSION + un-flip, on the device."""
    from sigma_b200 import _lib, fused
    torch.backends.cuda.matmul.allow_tf32 = False
    L = H * W
    K = 4 if kind == "cross4" else 2
    Ls = 2 * L if kind == "seq2" else L
    Cp = _lib.lib().sigma_ss2d_padded_cp(N, R)
    tag = f"extf/{kind}/{B}/{H}/{W}/{D}/{N}/{R}"
    xc = P.randn(S, tag + "/xc", (B, Ls, D)).cuda()
    xdbl = P.randn(S, tag + "/xdbl", (B, Ls, K, Cp))
    xdbl[..., 2 * N + R:] = 0.0
    xdbl = xdbl.cuda()
    dtw = P.rand(S, tag + "/dtw", (K, D, R), -R ** -0.5, R ** -0.5).cuda()
    dtb = P.rand(S, tag + "/dtb", (K, D), -6.0, -1.0).cuda()
    A = (-P.rand(S, tag + "/A", (K * D, N), 0.3, N + 0.5)).cuda()
    Ds = P.randn(S, tag + "/Ds", (K * D,)).cuda()
    idx = [torch.from_numpy(ix.copy()).cuda() for ix in _dir_index(kind, H, W)]
    us = torch.cat([xc[:, ix].transpose(1, 2) for ix in idx], 1).contiguous()
    dts = torch.cat([torch.einsum("blr,dr->bdl", xdbl[:, ix, k, 2 * N:2 * N + R], dtw[k]) for k, ix in enumerate(idx)], 1).contiguous()
    Bs = torch.stack([xdbl[:, ix, k, 0:N].transpose(1, 2) for k, ix in enumerate(idx)], 1).contiguous()
    Cs = torch.stack([xdbl[:, ix, k, N:2 * N].transpose(1, 2) for k, ix in enumerate(idx)], 1).contiguous()
    o, _ = ext.fwd(us, dts, A, Bs, Cs, Ds, dtb.reshape(-1).contiguous(), True, 1)
    ref = torch.zeros((K, B, Ls, D), device="cuda")
    for k, ix in enumerate(idx):
        ref[k][:, ix] = o[:, k * D:(k + 1) * D].transpose(1, 2)
    kid = {"cross4": _lib.DIRS_CROSS4, "seq2": _lib.DIRS_SEQ2}[kind]
    y = fused.ss2d_scan(kid, xc, xdbl, dtw, dtb, A, Ds, B, H, W, D, N, R, Cp)
    scale = float(ref.abs().max())
    err = float((y - ref).abs().max()) / scale
    assert err <= 1e-3, f"fused scan vs extension: {err:.3e} of the output scale"
