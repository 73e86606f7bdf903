"""Small drivers for `ncu --set full` captures of the round-2 kernels (one or two launches each).
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2 -o gpurun_out/<name> python scripts/ncu_targets.py <which>
which: opfwd (sigma_scan_fwd enc1 B=8 fp32) | opfwd_n4 (dec1) | opbwd (sigma_scan_bwd enc1 B=8) | gemm (stage-2 in_proj, tf32 and tf32x3) |
       conv (CAB conv3x3 96->32 at 120x160, B=16) | fusedbwd (FusedSS2DCore forward + backward, enc1 shape)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import fused, ops  # noqa: E402

which = sys.argv[1]
g = torch.Generator(device="cuda").manual_seed(1)


def scan_inputs(B, KD, L, N, K):
    u = torch.randn(B, KD, L, device="cuda", generator=g)
    dl = torch.randn(B, KD, L, device="cuda", generator=g) * 0.7
    A = -(torch.rand(KD, N, device="cuda", generator=g) * N + 0.3)
    Bm = torch.randn(B, K, N, L, device="cuda", generator=g)
    Cm = torch.randn(B, K, N, L, device="cuda", generator=g)
    D = torch.randn(KD, device="cuda", generator=g)
    bias = torch.rand(KD, device="cuda", generator=g) * 4 - 6
    return u, dl, A, Bm, Cm, D, bias


if which in ("opfwd", "opfwd_n4"):
    a = scan_inputs(8, 1536, 4800, 16 if which == "opfwd" else 4, 4)
    for _ in range(2):
        ops.selective_scan_cuda_core_fwd(*a, True, 1)
elif which == "opfwd_big":
    a = scan_inputs(32, 3072, 1200, 16, 4)
    for _ in range(2):
        ops.selective_scan_cuda_core_fwd(*a, True, 1)
elif which == "opbwd":
    a = scan_inputs(8, 1536, 4800, 16, 4)
    dout = torch.randn(8, 1536, 4800, device="cuda", generator=g)
    for _ in range(2):
        ops.selective_scan_cuda_core_bwd(*a, dout, None, True, 1)
elif which == "gemm":
    A = torch.randn(74 * 2 * 1200, 384, device="cuda", generator=g)
    W = torch.randn(1536, 384, device="cuda", generator=g) * 384 ** -0.5
    for tf32 in (True, False):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(2):
            fused.linear(A, W)
elif which == "conv":
    conv = torch.nn.Conv2d(96, 32, 3, 1, 1).cuda()
    x = torch.randn(16, 120, 160, 96, device="cuda", generator=g)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(2):
            fused.conv3x3(x, conv, gelu=True)
elif which == "fusedbwd":      # the fused core's forward (SAVE build) + reverse sweep on the enc1 shape, 4 streams
    from sigma_b200 import _lib
    B, H, W, D, N, R = 4, 60, 80, 384, 16, 12
    xc = torch.randn(B, H * W, D, device="cuda", generator=g).requires_grad_(True)
    xpw = (torch.randn(4, R + 2 * N, D, device="cuda", generator=g) * D ** -0.5).requires_grad_(True)
    dtw = ((torch.rand(4, D, R, device="cuda", generator=g) * 2 - 1) * R ** -0.5).requires_grad_(True)
    dtb = (torch.rand(4, D, device="cuda", generator=g) * 4 - 5).requires_grad_(True)
    Al = torch.log(torch.rand(4 * D, N, device="cuda", generator=g) * N + 0.5).requires_grad_(True)
    Ds = torch.randn(4 * D, device="cuda", generator=g).requires_grad_(True)
    for _ in range(2):
        y = ops.FusedSS2DCore.apply(xc, xpw, dtw, dtb, Al, Ds, _lib.DIRS_CROSS4, H, W)
        y.square().sum().backward()
torch.cuda.synchronize()
print("done", which)
