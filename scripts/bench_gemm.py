"""Dense-projection GEMMs of Sigma-tiny at `--images` per GPU: our tcgen05 TF32 kernel vs cuBLAS TF32 (torch.mm).
Reports ms and effective GB/s over (A read once + C written once [+ residual read])."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma_b200 import fused  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=16)
ap.add_argument("--only", nargs="*", default=None)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = True
S = 2 * a.images
# name, M, N, K, residual
SHAPES = [
    ("in_proj0", S * 19200, 384, 96, False), ("x_proj0", S * 19200, 160, 192, False), ("out_proj0", S * 19200, 96, 192, True),
    ("in_proj1", S * 4800, 768, 192, False), ("x_proj1", S * 4800, 176, 384, False), ("out_proj1", S * 4800, 192, 384, True),
    ("in_proj2", S * 1200, 1536, 384, False), ("x_proj2", S * 1200, 224, 768, False), ("out_proj2", S * 1200, 384, 768, True),
    ("in_proj3", S * 300, 3072, 768, False), ("x_proj3", S * 300, 320, 1536, False), ("out_proj3", S * 300, 768, 1536, True),
    ("merge0", S * 4800, 192, 384, False), ("dec_x_proj", a.images * 19200, 64, 192, False),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, M, N, K, res in SHAPES:
    if a.only and name not in a.only:
        continue
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * K ** -0.5
    R = torch.randn(M, N, device="cuda") if res else None
    out = torch.empty(M, N, device="cuda")
    byt = 4 * (M * K + M * N * (2 if res else 1) + N * K)
    line = f"{name:11s} M={M:7d} N={N:5d} K={K:5d}: "
    for mode in ("tcgen05", "cublas"):
        fused.USE_TCGEN05_GEMM = mode == "tcgen05"
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fused.linear(A, W, None, out=out, residual=R)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts[1:])[2]
        line += f"{mode} {ms:7.3f} ms {byt / ms / 1e6:7.1f} GB/s   "
    print(line, flush=True)
