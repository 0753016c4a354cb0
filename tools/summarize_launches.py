"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time per kernel name, share of the step."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v * scale))
tot = sum(t for _, t in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, t in rows:
    n = re.sub(r"\(.*", "", n)
    agg[n][0] += 1
    agg[n][1] += t
print("total %.1f us over %d launches" % (tot, len(rows)))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us  %5.1f%%  x%-4d %s" % (t, 100 * t / max(tot, 1e-9), c, n[:110]))
