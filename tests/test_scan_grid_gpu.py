"""GPU: the reference's own parity grid (models/encoders/selective_scan/test_selective_scan.py:137-224) reproduced in full —
itype {fp32, fp16, bf16} x seqlen {64 ... 4096} x delta_bias x softplus x D x groups {1 (3-D B/C), 2} x nrows {1..4} = 1 920
combinations, forward AND all seven gradients — through the same call the reference test makes (selective_scan_fn), with the
reference's tolerances.  The checker is the C oracle (oracle/selective_scan_ref.c, pinned to the reference's
selective_scan_ref + autograd goldens in tests/test_oracle.py) fed the SAME rounded 16-bit inputs, where the reference test
runs selective_scan_ref on the GPU.  One pytest case per (itype, seqlen) loops over the 64 inner combinations."""
import itertools

import numpy as np
import pytest
import torch

from oracle import scan_oracle

pytestmark = pytest.mark.gpu

SEQLENS = [64, 128, 256, 372, 512, 784, 1024, 1134, 2048, 4096]
ITYPES = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def _close(got, ref, rtol, atol, what):
    """torch.allclose with the reference's tolerances, plus a float32-resolution floor relative to the tensor's largest
    element: the reference's absolute tolerances assume O(1..100) gradients, but dA / ddelta_bias reach 1e5 at L = 4096,
    where 1e-5 of the largest value is the accumulated fp32 rounding of ANY implementation (the reference's own GPU
    comparison is fp32 against fp32; our checker accumulates in double)."""
    got = got.detach().float().cpu().numpy()
    atol = max(atol, 1e-5 * float(np.abs(ref).max()))
    bad = np.abs(got - ref) > atol + rtol * np.abs(ref)
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.size} out of tolerance, max abs err {np.abs(got - ref).max():.3e}"


@pytest.mark.parametrize("seqlen", SEQLENS)
@pytest.mark.parametrize("iname", list(ITYPES))
def test_reference_grid(iname, seqlen):
    from sigma_b200 import ops
    itype = ITYPES[iname]
    rtol, atol = (6e-4, 2e-3) if itype == torch.float32 else (3e-3, 5e-3)     # test_selective_scan.py:148-150
    if itype == torch.bfloat16:
        rtol, atol = 3e-2, 5e-2
    rtolw, atolw = 1e-3, 1e-3
    batch, dim, dstate = 2, 24, 8
    for groups, has_D, has_bias, softplus, nrows in itertools.product([1, 2], [False, True], [False, True], [False, True], [1, 2, 3, 4]):
        torch.manual_seed(0)
        dev = "cuda"
        A = (-0.5 * torch.rand(dim, dstate, device=dev)).requires_grad_()
        bshape = (batch, dstate, seqlen) if groups == 1 else (batch, groups, dstate, seqlen)
        B = torch.randn(*bshape, device=dev, dtype=itype, requires_grad=True)
        C = torch.randn(*bshape, device=dev, dtype=itype, requires_grad=True)
        D = torch.randn(dim, device=dev, requires_grad=True) if has_D else None
        bias = (0.5 * torch.rand(dim, device=dev)).requires_grad_() if has_bias else None
        u = torch.randn(batch, dim, seqlen, device=dev, dtype=itype, requires_grad=True)
        delta = (0.5 * torch.rand(batch, dim, seqlen, device=dev, dtype=itype)).requires_grad_()
        out = ops.selective_scan_fn(u, delta, A, B, C, D, delta_bias=bias, delta_softplus=softplus, nrows=nrows)
        g = torch.randn_like(out)
        out.backward(g)
        f = lambda t: None if t is None else t.detach().float().cpu().numpy()
        B4 = f(B) if groups > 1 else f(B)[:, None]
        C4 = f(C) if groups > 1 else f(C)[:, None]
        ref = scan_oracle.scan_fwd(f(u), f(delta), f(A), B4, C4, f(D), f(bias), softplus)
        what = f"{iname} L={seqlen} g={groups} D={has_D} bias={has_bias} sp={softplus} nrows={nrows}"
        assert out.dtype == itype
        _close(out, ref, rtol, atol, what + " out")
        du, dd, dA, dB, dC, dD, dbias = scan_oracle.scan_bwd(f(u), f(delta), f(A), B4, C4, f(D), f(bias), f(g), softplus)
        assert u.grad.dtype == itype and B.grad.dtype == itype and B.grad.shape == B.shape
        _close(u.grad, du, rtol * 2, atol * 2, what + " du")                    # :216-224
        _close(delta.grad, dd, rtol * 5, atol * 10, what + " ddelta")
        _close(A.grad, dA, rtolw, atolw * 5, what + " dA")
        _close(B.grad.reshape(B4.shape), dB, rtol, atol, what + " dB")
        _close(C.grad.reshape(C4.shape), dC, rtol, atol, what + " dC")
        if has_D:
            _close(D.grad, dD, rtolw, atolw, what + " dD")
        if has_bias:
            _close(bias.grad, dbias, rtolw, atolw, what + " dbias")
