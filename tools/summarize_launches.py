"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    agg[name][0] += 1
    agg[name][1] += v
    total += v
print(f"| kernel | launches | total us | mean us | share |")
print("|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {t/1e3:.1f} | {t/1e3/n:.2f} | {100*t/total:.1f}% |")
print(f"| total | {sum(v[0] for v in agg.values())} | {total/1e3:.1f} | | 100% |")
