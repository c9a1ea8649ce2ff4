"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel-name count, total time, share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], v * scale))
agg = defaultdict(lambda: [0, 0.0])
for name, us in rows:
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)[:90]
    agg[name][0] += 1
    agg[name][1] += us
total = sum(v[1] for v in agg.values())
print(f"launches {len(rows)}  total {total/1000:.3f} ms")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{us/1000:9.3f} ms {100*us/total:5.1f}%  x{n:<5d} {name}")
