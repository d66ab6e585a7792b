"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0.0, 0])
total = 0.0
n = 0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("fnp::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name][0] += us
    tot[name][1] += 1
    total += us
    n += 1
print(f"# launch list summary: {path}\n")
print(f"{n} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES)\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for name, (us, k) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{name}` | {k} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / k:.1f} |")
