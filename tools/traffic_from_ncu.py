"""DRAM traffic of the tcgen05 launches of one bench step, per roofline class.
Inputs: an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"gemm_tcgen05|wgrad_halo" --csv` log of
`bench.py --steps 1 --warmup 1` and the PASSL_B200_BENCH_LAUNCH_TABLE file the same run wrote (one line per launch of the
instrumented step, in launch order, with its class).  The last len(table) gemm launches of the capture are the instrumented
step.  Writes / updates profiles/r02_traffic.json entries '<cfg>_hbm' and '<cfg>_tensor' (bytes per launch, class average).
usage: traffic_from_ncu.py ncu.csv table.txt cfg [traffic.json]"""
import csv
import json
import os
import re
import sys

ncu_csv, table, cfg = sys.argv[1], sys.argv[2], sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
rows = []
for l in open(table):
    m = re.match(r"\s*([\d.]+) us\s+([\d.]+) of (\w+)\s+roofline\s+([\d.]+) TF/s\s+([\d.]+) GB/s\s+(.*)", l)
    rows.append(dict(us=float(m[1]), cls=m[3], alg=float(m[5]) * 1e9 * float(m[1]) * 1e-6, desc=m[6]))
with open(ncu_csv) as f:
    lines = [l for l in f if l.startswith('"')]
per = {}
order = []
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in csv.DictReader(lines):
    if "gemm_tcgen05" not in r["Kernel Name"] and "wgrad_halo" not in r["Kernel Name"]:
        continue
    i = r["ID"]
    if i not in per:
        per[i] = 0.0
        order.append(i)
    if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        per[i] += float(r["Metric Value"].replace(",", "")) * scale.get(r["Metric Unit"], 1.0)
ids = order[-len(rows):]
assert len(ids) == len(rows), (len(ids), len(rows))
db = json.load(open(out)) if os.path.exists(out) else {}
for cls in ("hbm", "tensor"):
    sel = [(per[i], r) for i, r in zip(ids, rows) if r["cls"] == cls]
    if not sel:
        continue
    dram = sum(d for d, _ in sel)
    alg = sum(r["alg"] for _, r in sel)
    db["%s_%s" % (cfg, cls)] = {
        "dram_bytes": dram / len(sel), "algorithmic_bytes": alg / len(sel), "ratio": dram / alg,
        "launch": "average over the %d %s-class tcgen05 launches of one step" % (len(sel), cls),
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of every gemm_tcgen05 / wgrad_halo launch of one bench step "
                  "(%s), paired in launch order with the step's launch table" % os.path.basename(ncu_csv)}
    print(cls, "launches", len(sel), "dram MB/launch %.1f" % (dram / len(sel) / 1e6), "algorithmic %.1f" % (alg / len(sel) / 1e6),
          "ratio %.3f" % (dram / alg))
json.dump(db, open(out, "w"), indent=1)
