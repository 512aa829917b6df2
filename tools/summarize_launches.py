"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    if not (len(sys.argv) > 2 and sys.argv[2] == "full"):
        name = re.sub(r"<.*", "", name)
    name = name.replace("bk::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1e-9)
    tot[name][0] += 1
    tot[name][1] += v * scale
total = sum(v[1] for v in tot.values())
print(f"{'kernel':48s} {'launches':>9s} {'seconds':>10s} {'share':>7s}")
for k, (c, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:48]:48s} {c:9d} {s:10.4f} {100 * s / total:6.1f}%")
print(f"{'TOTAL':48s} {sum(v[0] for v in tot.values()):9d} {total:10.4f}")
