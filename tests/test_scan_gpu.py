"""GPU: op-level parity of the sm_100a selective scan through the C-ABI (sigma_scan_fwd), against
(a) the committed goldens of the reference's selective_scan_ref and (b) the CPU oracle on seeded
inputs.  Tolerances are the reference's own (test_selective_scan.py:148-151): fp32 rtol 6e-4 /
atol 2e-3, fp16 3e-3/5e-3, bf16 3e-2/5e-2 — and BASELINE.json's 1e-3 fp32 / 1e-2 bf16 relative to
the output scale."""
import glob
import os

import numpy as np
import pytest
import torch

import procedural as P
from helpers import GOLDEN, SEED, assert_close
from oracle import scan_oracle

pytestmark = pytest.mark.gpu


def _run(u, dl, A, Bm, Cm, D, bias, sp, dtype=torch.float32, split=0, nrows=1):
    from sigma_b200 import ops
    dev = "cuda"
    c = lambda t: None if t is None else t.to(dev)
    out, x = ops.selective_scan_cuda_core_fwd(c(u).to(dtype), c(dl).to(dtype), c(A), c(Bm).to(dtype), c(Cm).to(dtype),
                                              c(D), c(bias), sp, nrows, _force_split=split)
    torch.cuda.synchronize()
    return out, x


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "scan_case*.npz"))) +
                         [os.path.join(GOLDEN, "scan_config1.npz")], ids=os.path.basename)
@pytest.mark.parametrize("split", [0, 1, 3])
def test_fwd_matches_reference_golden(path, split):
    g = np.load(path)
    b, d, n, L, G, hD, hb, sp = (int(v) for v in g["cfg"])
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED, b, d, n, L, G, has_D=bool(hD), has_bias=bool(hb))
    out, x = _run(u, dl, A, Bm, Cm, D, bias, sp, split=split)
    assert out.shape == (b, d, L) and x.shape == (b, d, (L + 2047) // 2048, 2 * n)
    if "out_sub" in g.files:
        ref, got = g["out_sub"], out[:, ::16]
    else:
        ref, got = g["out"], out
    assert_close(got, ref, 6e-4, 2e-3, os.path.basename(path))
    scale = float(np.abs(ref).max())
    assert float(np.abs(got.cpu().numpy() - ref).max()) < 1e-3 * scale      # BASELINE.json: 1e-3 fp32


@pytest.mark.parametrize("dtype,rtol,atol", [(torch.float16, 3e-3, 5e-3), (torch.bfloat16, 3e-2, 5e-2)])
def test_fwd_half_io(dtype, rtol, atol):
    b, d, n, L, G = 2, 24, 8, 372, 2
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED, b, d, n, L, G)
    out, _ = _run(u, dl, A, Bm, Cm, D, bias, True, dtype=dtype)
    assert out.dtype == dtype
    q = lambda t: t.to(dtype).float().numpy()
    ref = scan_oracle.scan_fwd(q(u), q(dl), A.numpy(), q(Bm), q(Cm), D.numpy(), bias.numpy(), True)
    assert_close(out, ref, rtol, atol, str(dtype))


@pytest.mark.parametrize("b,d,n,L,G", [
    (1, 1, 1, 1, 1), (1, 5, 3, 7, 1), (2, 33, 16, 31, 1), (2, 64, 16, 33, 2), (1, 96, 16, 300, 4),
    (2, 40, 4, 1200, 2), (1, 8, 32, 129, 1), (1, 4, 64, 70, 1), (1, 768, 16, 2400, 4), (1, 36, 8, 4801, 3),
])
def test_fwd_shapes_and_edges(b, d, n, L, G):
    """ragged L (not a multiple of 4/32/2048), dim not a multiple of 32, groups, wide dstate."""
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED + 3, b, d, n, L, G)
    ref = scan_oracle.scan_fwd(u.numpy(), dl.numpy(), A.numpy(), Bm.numpy(), Cm.numpy(), D.numpy(), bias.numpy(), True)
    for split in (0, 2):
        out, x = _run(u, dl, A, Bm, Cm, D, bias, True, split=split)
        assert_close(out, ref, 6e-4, 2e-3, f"{(b, d, n, L, G)} split={split}")


def test_fwd_strided_inputs_and_x_states():
    """non-contiguous batch/dim strides (the boundary passes element strides, selective_scan.h:27) and the
    chunk-end states x = (prod a, h) every 2048 positions (fwd_kernel.cuh:181-184)."""
    from sigma_b200 import ops
    b, d, n, L, G = 2, 16, 4, 4500, 1
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED + 4, b, d, n, L, G)
    big_u = torch.zeros(b, d * 2, L + 8, device="cuda")
    big_u[:, ::2, 3:L + 3] = u.cuda()
    u_view = big_u[:, ::2, 3:L + 3]                     # strides (2d(L+8), 2(L+8), 1), misaligned base
    assert u_view.stride(-1) == 1 and not u_view.is_contiguous()
    out, x = ops.selective_scan_cuda_core_fwd(u_view, dl.cuda(), A.cuda(), Bm.cuda(), Cm.cuda(), D.cuda(), bias.cuda(), True, 1)
    ref = scan_oracle.scan_fwd(u.numpy(), dl.numpy(), A.numpy(), Bm.numpy(), Cm.numpy(), D.numpy(), bias.numpy(), True)
    assert_close(out, ref, 6e-4, 2e-3, "strided u")
    # x: re-derive h at the end of each chunk with D=None, C = one-hot readout of state n
    for c, lend in enumerate([2048, 4096, 4500]):
        for s in range(n):
            Csel = torch.zeros_like(Cm)
            Csel[:, :, s] = 1.0
            h = scan_oracle.scan_fwd(u[:, :, :lend].numpy(), dl[:, :, :lend].numpy(), A.numpy(), Bm[..., :lend].numpy(),
                                     Csel[..., :lend].numpy(), None, bias.numpy(), True)[:, :, -1]
            assert_close(x[:, :, c, 2 * s + 1], h, 6e-4, 2e-3, f"x.h chunk{c} state{s}")
        # running prefix since the SEQUENCE start (SSMScanPrefixCallbackOp, fwd_kernel.cuh:181-184), not per chunk
        dls = torch.nn.functional.softplus(dl[:, :, :lend] + bias[None, :, None]).double().sum(-1)
        prod_a = torch.exp(dls[..., None] * A.double()[None])
        assert_close(x[:, :, c, 0::2], prod_a.float(), 2e-3, 1e-6, f"x.prod_a chunk{c}")


def test_errors_are_loud():
    from sigma_b200 import ops
    u = torch.zeros(1, 6, 8, device="cuda")
    A = torch.zeros(6, 4, device="cuda")
    Bm = torch.zeros(1, 4, 4, 8, device="cuda")           # 6 % 4 != 0
    with pytest.raises(RuntimeError):
        ops.selective_scan_cuda_core_fwd(u, u, A, Bm, Bm, None, None, False, 1)
    with pytest.raises(RuntimeError):                     # CPU tensors: no fallback
        ops.selective_scan_cuda_core_fwd(u.cpu(), u.cpu(), A.cpu(), Bm.cpu(), Bm.cpu(), None, None, False, 1)


@pytest.mark.parametrize("b,d,n,L,G", [(2, 64, 16, 300, 2), (1, 192, 16, 4800, 1), (3, 96, 4, 1204, 3), (1, 256, 8, 2400, 4),
                                       (2, 768, 16, 1200, 4), (1, 128, 4, 4500, 1)])
@pytest.mark.parametrize("dn", ["f32", "bf16", "f16"])
def test_fwd_tma_path(b, d, n, L, G, dn):
    """Shapes the TMA-staged kernel takes (channel groups % 32 == 0, 16-byte rows): all element types natively, 1 / 3 / 7
    L-segments, chunk states with the running prefix, and agreement with the generic kernel (SIGMA_OP_GENERIC)."""
    import os
    from sigma_b200 import ops
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[dn]
    if dn != "f32" and L % 8:
        L = L - L % 8
    u, dl, A, Bm, Cm, D, bias = P.scan_inputs(SEED + 9, b, d, n, L, G)
    q = lambda t: t.to(dt)
    f = lambda t: q(t).float().numpy()
    ref = scan_oracle.scan_fwd(f(u), f(dl), A.numpy(), f(Bm), f(Cm), D.numpy(), bias.numpy(), True)
    rt, at = (6e-4, 2e-3) if dn == "f32" else ((3e-3, 5e-3) if dn == "f16" else (3e-2, 5e-2))
    args = (q(u).cuda(), q(dl).cuda(), A.cuda(), q(Bm).cuda(), q(Cm).cuda(), D.cuda(), bias.cuda(), True, 1)
    out, x = ops.selective_scan_cuda_core_fwd(*args)
    assert out.dtype == dt
    assert_close(out, ref, rt, at, f"tma {dn} {(b, d, n, L, G)}")
    if dn == "f32":
        for split in (3, 7):
            o2, x2 = ops.selective_scan_cuda_core_fwd(*args, _force_split=split)
            assert_close(o2, ref, rt, at, f"tma split={split}")
            assert_close(x2, x.cpu().numpy(), 2e-3, 1e-5, f"x states split={split}")
    os.environ["SIGMA_OP_GENERIC"] = "1"
    try:
        og, xg = ops.selective_scan_cuda_core_fwd(*args)
    finally:
        del os.environ["SIGMA_OP_GENERIC"]
    assert_close(out, og.float().cpu().numpy(), rt, at, "tma vs generic kernel")
    assert_close(x, xg.cpu().numpy(), 2e-3, 1e-5, "x states tma vs generic")
