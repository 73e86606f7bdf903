"""Summarise .ncu-rep captures (ncu --set full) into a few lines per launch: python scripts/ncu_rep_summary.py rep1 rep2 ... > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"## {rep}: unreadable")
        continue
    hdr, units = rows[0], rows[1]
    print(f"## {rep}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"{d.get('Kernel Name', '')[:100]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"    {k:75s} {d[k]} {u.get(k, '')}")
        st = sorted(((float(v), k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')) for k, v in d.items()
                     if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k and v not in ('', None)), reverse=True)
        print("    warp stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
