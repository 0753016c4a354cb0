"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time per kernel name, share of the step.
With --marker NAME only the launches from the first kernel whose name contains NAME up to (excluding) the next one are counted
(= exactly one training step when NAME is the first kernel of the step, e.g. stem_pack_input)."""
import csv
import re
import sys
from collections import defaultdict

args = sys.argv[1:]
marker = None
if "--marker" in args:
    i = args.index("--marker")
    marker = args[i + 1]
    del args[i:i + 2]
rows = []
with open(args[0]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v * scale))
if marker:
    idx = [i for i, (n, _) in enumerate(rows) if marker in n]
    if len(idx) >= 2:
        rows = rows[idx[0]:idx[1]]
        print("one step: launches %d..%d of the capture (between two '%s')" % (idx[0], idx[1] - 1, marker))
    else:
        print("marker '%s' found %d times: summarising the whole capture" % (marker, len(idx)))
tot = sum(t for _, t in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, t in rows:
    n = re.sub(r"\(.*", "", n)
    agg[n][0] += 1
    agg[n][1] += t
print("total %.1f us over %d launches (ncu per-launch times: cold cache, serialised — use the SHARES)" % (tot, len(rows)))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us  %5.1f%%  x%-4d %s" % (t, 100 * t / max(tot, 1e-9), c, n[:110]))
