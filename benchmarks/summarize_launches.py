"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: time per kernel name, share of the step."""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
tot = sum(v for _, v in rows)
agg = defaultdict(lambda: [0, 0.0])
for k, v in rows:
    name = re.sub(r"\(.*", "", k)[:90]
    agg[name][0] += 1; agg[name][1] += v
print(f"{len(rows)} launches, {tot:.1f} us total (serialised, cold cache: compare shares)")
for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:9.1f} us {100*v/tot:5.1f}%  x{n:<3d} {name}")
