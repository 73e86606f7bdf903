"""Op-level scan (sigma_scan_fwd / sigma_scan_bwd): time vs the number of L-segments, to set the split policy.
    python scripts/bench_op_splits.py [--batch 1 2 8] [--bwd]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sigma_b200 import ops  # noqa: E402

SHAPES = [("enc0", 768, 19200, 16, 4), ("enc1", 1536, 4800, 16, 4), ("enc2", 3072, 1200, 16, 4), ("dec0", 768, 19200, 4, 4),
          ("dec1", 1536, 4800, 4, 4), ("cromb1", 384, 4800, 4, 1), ("conmb0", 384, 38400, 4, 2)]


def timeit(fn, flush, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 2, 8])
    ap.add_argument("--splits", type=int, nargs="+", default=[0, 1, 2, 4, 8, 16, 32])
    ap.add_argument("--bwd", action="store_true")
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    for B in a.batch:
        for name, KD, L, N, K in SHAPES:
            u = torch.randn(B, KD, L, device="cuda", generator=g)
            dl = torch.randn(B, KD, L, device="cuda", generator=g) * 0.7
            A = -(torch.rand(KD, N, device="cuda", generator=g) * N + 0.3)
            Bm = torch.randn(B, K, N, L, device="cuda", generator=g)
            Cm = torch.randn(B, K, N, L, device="cuda", generator=g)
            D = torch.randn(KD, device="cuda", generator=g)
            bias = torch.rand(KD, device="cuda", generator=g) * 4 - 6
            dout = torch.randn(B, KD, L, device="cuda", generator=g)
            row = []
            for sp in a.splits:
                if a.bwd:
                    t = timeit(lambda: ops.selective_scan_cuda_core_bwd(u, dl, A, Bm, Cm, D, bias, dout, None, True, 1, _force_split=sp), flush, 3)
                else:
                    t = timeit(lambda: ops.selective_scan_cuda_core_fwd(u, dl, A, Bm, Cm, D, bias, True, 1, _force_split=sp), flush)
                row.append(f"{'auto' if sp == 0 else sp}:{t:.3f}")
            print(f"{'bwd' if a.bwd else 'fwd'} B={B} {name:7s} N={N:2d} warps={B * KD // 32:5d}  " + "  ".join(row), flush=True)
            del u, dl, Bm, Cm, dout


if __name__ == "__main__":
    main()
