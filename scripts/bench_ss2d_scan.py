"""Micro-benchmark of the fused SS2D scan (sigma_ss2d_scan_fwd) on the Sigma-tiny call shapes.
    python scripts/bench_ss2d_scan.py [--images 16] [--only enc0 dec0] [--split 0]
Reports ms, algorithmic GB/s (SURVEY.md §8d formula), and the exp rate (G ex2/s; MUFU peak = 16/clk/SM)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import _lib, fused  # noqa: E402

# name, kind, streams-per-image, H, W, D, N, R
SHAPES = [
    ("enc0", "cross4", 2, 120, 160, 192, 16, 6), ("enc1", "cross4", 2, 60, 80, 384, 16, 12),
    ("enc2", "cross4", 2, 30, 40, 768, 16, 24), ("enc3", "cross4", 2, 15, 20, 1536, 16, 48),
    ("dec0", "cross4", 1, 120, 160, 192, 4, 6), ("dec1", "cross4", 1, 60, 80, 384, 4, 12), ("dec2", "cross4", 1, 30, 40, 768, 4, 24),
    ("conmb0", "seq2", 1, 120, 160, 192, 4, 6), ("cromb0", "cross", 2, 120, 160, 192, 4, 6),
    ("conmb1", "seq2", 1, 60, 80, 384, 4, 12), ("cromb1", "cross", 2, 60, 80, 384, 4, 12),
    ("conmb2", "seq2", 1, 30, 40, 768, 4, 24), ("cromb2", "cross", 2, 30, 40, 768, 4, 24),
    ("conmb3", "seq2", 1, 15, 20, 1536, 4, 48), ("cromb3", "cross", 2, 15, 20, 1536, 4, 48),
]
KID = {"cross4": _lib.DIRS_CROSS4, "seq2": _lib.DIRS_SEQ2, "cross": _lib.DIRS_CROSS}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--split", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fused._FORCE_SPLIT = a.split
    for name, kind, spi, H, W, D, N, R in SHAPES:
        if a.only and name not in a.only:
            continue
        Bt = a.images * spi
        L = H * W
        Kx = {"cross4": 4, "seq2": 2, "cross": 1}[kind]
        Kw = 2 if kind == "cross" else Kx
        Ls = 2 * L if kind == "seq2" else L
        Cp = _lib.lib().sigma_ss2d_padded_cp(N, R)
        xc = torch.randn(Bt, Ls, D, device="cuda")
        xdbl = torch.randn(Bt, Ls, Kx, Cp, device="cuda")
        dtw = (torch.rand(Kw, D, R, device="cuda") * 2 - 1) * R ** -0.5
        dtb = torch.rand(Kw, D, device="cuda") * 5 - 6
        A = -(torch.rand(Kw * D, N, device="cuda") * N + 0.5)
        Ds = torch.randn(Kw * D, device="cuda")
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fused.ss2d_scan(KID[kind], xc, xdbl, dtw, dtb, A, Ds, Bt, H, W, D, N, R, Cp)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts[1:])[len(ts[1:]) // 2]
        KD = Kx * D
        byt = 4 * (3 * Bt * KD * Ls + 2 * Bt * Kx * N * Ls) + 4 * (KD * N + 2 * KD)
        exps = Bt * KD * Ls * N
        print(f"{name:7s} images={a.images:3d} streams={Bt:3d} D={D:5d} L={Ls:6d} N={N:2d} R={R:2d}: {ms:8.3f} ms  "
              f"{byt / ms / 1e6:8.1f} GB/s (frac {byt / ms / 1e6 / peak:.3f})  {exps / ms / 1e6:7.1f} Gex2/s", flush=True)
        del xc, xdbl


if __name__ == "__main__":
    main()
