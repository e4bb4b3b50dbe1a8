"""Summarise an ncu launch list (gpu__time_duration.sum csv): per-kernel totals for the LAST step.
usage: python tools/summarize_launches.py file.csv [n_steps]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
import gzip
with (gzip.open(path, "rt") if path.endswith(".gz") else open(path)) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    grid = r.get("Grid Size", "")
    rows.append((r["Kernel Name"], us, grid))
n = len(rows) // nsteps
last = rows[-n:]
tot = sum(u for _, u, _ in last)
agg = defaultdict(lambda: [0.0, 0])
for k, u, g in last:
    k = re.sub(r"\(.*", "", k)
    agg[k][0] += u
    agg[k][1] += 1
print(f"launches in last step: {n}, total kernel time {tot/1e3:.3f} ms")
for k, (u, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{u/1e3:9.3f} ms {100*u/tot:5.1f}%  x{c:4d}  {k[:110]}")
if len(sys.argv) > 3:
    print("--- top individual launches")
    for k, u, g in sorted(last, key=lambda x: -x[1])[:25]:
        print(f"{u:9.1f} us  grid {g:>18s}  {re.sub(r'[(].*', '', k)[:90]}")
