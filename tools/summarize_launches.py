"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void ", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
    tot[name][0] += 1
    tot[name][1] += v * scale
total = sum(v[1] for v in tot.values())
print(f"| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:90]}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
print(f"| **total** | {sum(v[0] for v in tot.values())} | {total:.3f} | 100% |")
