"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9, "s": 1e9}.get(unit, 1)
    name = re.sub(r"<.*", "", r["Kernel Name"].split("(")[0])[:70]
    rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"# {len(rows)} launches, {tot / 1e6:.3f} ms total (serialised, cold-cache: compare shares)")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ns / tot * 100:6.2f}%  {ns / 1e6:9.3f} ms  {c:5d}x  {n}")
