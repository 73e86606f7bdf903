"""L-segment sweep of the fused SS2D scan at small batch (the latency regime): every Sigma-tiny call shape x images in {1, 2, 4, 8} x
forced segment counts, next to the library's own choice under the old rule ("fill 592 warp slots") and the current cost model.
    python scripts/bench_ss2d_splits.py > profiles/r02_ss2d_split_sweep.txt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import _lib, fused  # noqa: E402
from bench_ss2d_scan import SHAPES, KID  # noqa: E402

SPLITS = [1, 2, 3, 4, 6, 8, 12, 16, 24, 32]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=5):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[len(ts[1:]) // 2]


print("# times in microseconds; 'old' / 'new' = the library's own segment choice under SIGMA_SCAN_SPLIT_RULE=old / default")
print(f"{'shape':8s} {'img':>3s} " + " ".join(f"n={n:<4d}" for n in SPLITS) + "    old    new   best")
tot = {"old": 0.0, "new": 0.0, "best": 0.0}
for images in (1, 2, 4, 8):
    for name, kind, spi, H, W, D, N, R in SHAPES:
        Bt = images * spi
        Kx = {"cross4": 4, "seq2": 2, "cross": 1}[kind]
        Kw = 2 if kind == "cross" else Kx
        Ls = 2 * H * W if kind == "seq2" else H * W
        Cp = _lib.lib().sigma_ss2d_padded_cp(N, R)
        xc = torch.randn(Bt, Ls, D, device="cuda")
        xdbl = torch.randn(Bt, Ls, Kx, Cp, device="cuda")
        dtw = (torch.rand(Kw, D, R, device="cuda") * 2 - 1) * R ** -0.5
        dtb = torch.rand(Kw, D, device="cuda") * 5 - 6
        A = -(torch.rand(Kw * D, N, device="cuda") * N + 0.5)
        Ds = torch.randn(Kw * D, device="cuda")
        call = lambda: fused.ss2d_scan(KID[kind], xc, xdbl, dtw, dtb, A, Ds, Bt, H, W, D, N, R, Cp)
        row = []
        for n in SPLITS:
            fused._FORCE_SPLIT = n
            row.append(timed(call) * 1e3)
        fused._FORCE_SPLIT = 0
        os.environ["SIGMA_SCAN_SPLIT_RULE"] = "old"
        t_old = timed(call) * 1e3
        del os.environ["SIGMA_SCAN_SPLIT_RULE"]
        t_new = timed(call) * 1e3
        tot["old"] += t_old; tot["new"] += t_new; tot["best"] += min(row)
        print(f"{name:8s} {images:3d} " + " ".join(f"{t:6.0f}" for t in row) + f" {t_old:6.0f} {t_new:6.0f} {min(row):6.0f}", flush=True)
print(f"# sum over all rows: old rule {tot['old']:.0f} us, cost model {tot['new']:.0f} us, per-row best forced count {tot['best']:.0f} us")
