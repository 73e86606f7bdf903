"""gpurun_out/<ref_ext json> -> profiles/<name>.md: the reference CUDA extension (rebuilt for sm_100a) next to sigma_b200's
op-level kernels and fused scan, same box, same tensors.   python scripts/make_ref_ext_table.py IN.json OUT.md"""
import json
import sys

d = json.load(open(sys.argv[1]))
L = [f"# Reference `selective_scan_cuda_core` (rebuilt for sm_100a) vs sigma_b200 — {d['gpu']}, HBM peak {d['peak_GBps']} GB/s (measured)", "",
     "Same process, same tensors, CUDA events, 256 MiB L2 flush between repetitions, median of 5 (backward: 3).  GB/s = algorithmic bytes of",
     "SURVEY.md §8d / time.  `ref` = best of nrows 1 / 4.  `op` = `sigma_scan_fwd` / `sigma_scan_bwd` (reference layout, drop-in);",
     "`fused` = `sigma_ss2d_scan_fwd` on the channels-last formulation of the same call (fp32 only).  err = max |ours − ext| / max |ext|.", "",
     "| shape | B | dtype | ref fwd ms | ref GB/s | op fwd ms | op GB/s | op frac | op speed-up | fused ms | fused speed-up | ref bwd ms | op bwd ms | bwd speed-up | fwd err | max bwd err |",
     "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in d["rows"]:
    best = min(r["ref_fwd_ms"].values())
    be = r.get("bwd_err_vs_ext")
    L.append(f"| {r['shape']} ({r['KD']}x{r['L']}, N={r['N']}) | {r['batch']} | {r['dtype']} | {best:.3f} | {r['ref_fwd_GBps']:.0f} | {r['our_fwd_ms']:.3f} | "
             f"{r['our_fwd_GBps']:.0f} | {r['our_fwd_frac']:.3f} | {r['fwd_speedup']:.2f}x | {r.get('fused_ms', float('nan')):.3f} | "
             f"{r.get('fused_speedup', float('nan')):.2f}x | {r.get('ref_bwd_ms', float('nan')):.3f} | {r.get('our_bwd_ms', float('nan')):.3f} | "
             f"{r.get('bwd_speedup', float('nan')):.2f}x | {r['fwd_err_vs_ext']:.1e} | {max(be.values()) if be else float('nan'):.1e} |")
open(sys.argv[2], "w").write("\n".join(L).replace("nan", "–") + "\n")
print("wrote", sys.argv[2], len(d["rows"]), "rows")
