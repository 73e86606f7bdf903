"""ncu CSV (scripts/scan_step_once.py capture) -> profiles/r02_scan_traffic.json: DRAM bytes moved by the fused-scan launches
of one step (dram__bytes_read.sum + dram__bytes_write.sum) and their duration, per launch and in total.
    python scripts/ncu_scan_traffic.py gpurun_out/r02_scan_traffic.csv --batch 74"""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--batch", type=int, default=74)
ap.add_argument("--model", default="sigma_tiny")
ap.add_argument("--height", type=int, default=480)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_scan_traffic.json"))
a = ap.parse_args()
lines = [ln for ln in open(a.csv) if ln.startswith('"')]
rows = [r for r in csv.DictReader(lines) if "ss2d_scan_kernel" in r["Kernel Name"]]   # the CSV may hold every kernel of the step
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}
per = {}
for r in rows:
    k = r["ID"]
    d = per.setdefault(k, {"name": r["Kernel Name"], "read": 0.0, "write": 0.0, "time": 0.0})
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1)
    if r["Metric Name"] == "dram__bytes_read.sum":
        d["read"] = v
    elif r["Metric Name"] == "dram__bytes_write.sum":
        d["write"] = v
    elif r["Metric Name"] == "gpu__time_duration.sum":
        d["time"] = v
n = len(per)
tot = sum(d["read"] + d["write"] for d in per.values())
out = dict(batch=a.batch, model=a.model, height=a.height, width=a.width, launches=n, dram_bytes_total=int(tot),
           dram_bytes_per_launch=int(tot / max(n, 1)), dram_read_bytes=int(sum(d["read"] for d in per.values())),
           dram_write_bytes=int(sum(d["write"] for d in per.values())), kernel_seconds_under_ncu=sum(d["time"] for d in per.values()),
           source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every ss2d_scan_kernel launch of one eager step "
                  "(scripts/scan_step_once.py)", csv=os.path.basename(a.csv))
json.dump(out, open(a.out, "w"), indent=1)
print(json.dumps(out))
