"""Op-level timing of sigma_scan_fwd on the Sigma shapes: ms and algorithmic GB/s.
Run on the GPU box:  python scripts/bench_scan_op.py [--batch 1 8]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sigma_b200 import ops  # noqa: E402

SHAPES = [  # (name, KD, L, N, K)
    ("enc0", 768, 19200, 16, 4), ("enc1", 1536, 4800, 16, 4), ("enc2", 3072, 1200, 16, 4), ("enc3", 6144, 300, 16, 4),
    ("dec0", 768, 19200, 4, 4), ("dec2", 3072, 1200, 4, 4), ("conmb0", 384, 38400, 4, 2), ("cromb0", 192, 19200, 4, 1),
]


def algo_bytes(B, KD, L, N, K):
    return 4 * (3 * B * KD * L + 2 * B * K * N * L) + 4 * (KD * N + 2 * KD)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 8])
    ap.add_argument("--splits", type=int, nargs="+", default=[0, 1])
    args = ap.parse_args()
    peak = 6486.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for B in args.batch:
        for name, KD, L, N, K in SHAPES:
            u = torch.randn(B, KD, L, device="cuda")
            dl = torch.rand(B, KD, L, device="cuda") * 0.5
            A = -torch.rand(KD, N, device="cuda") * 4
            Bm = torch.randn(B, K, N, L, device="cuda")
            Cm = torch.randn(B, K, N, L, device="cuda")
            D = torch.randn(KD, device="cuda")
            bias = torch.rand(KD, device="cuda") - 4
            for split in args.splits:
                ts = []
                for it in range(5):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.selective_scan_cuda_core_fwd(u, dl, A, Bm, Cm, D, bias, True, 1, _force_split=split)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = sorted(ts[2:])[len(ts[2:]) // 2]
                gbs = algo_bytes(B, KD, L, N, K) / ms / 1e6
                print(f"B={B:2d} {name:7s} KD={KD:5d} L={L:6d} N={N:2d} split={'auto' if split == 0 else split}: "
                      f"{ms:8.3f} ms  {gbs:8.1f} GB/s  frac {gbs / peak:.3f}", flush=True)
            del u, dl, Bm, Cm


if __name__ == "__main__":
    main()
