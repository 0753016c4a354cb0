"""SASS rows of an .ncu-rep in address order with stall samples, condensed: runs of instructions are merged into regions separated
by 'anchor' opcodes (barrier waits, TMEM loads, branches).  usage: ncu_sass.py file.ncu-rep [min_samples_to_print_row] [--launch=N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
launch = [a for a in sys.argv if a.startswith("--launch=")]
sys.argv = [a for a in sys.argv if not a.startswith("--launch=")]
sel = ["--launch-skip", launch[0].split("=")[1], "--launch-count", "1"] if launch else []
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
acc, accn, start = 0, 0, None
tot = 0
for r in rows[2:]:
    if len(r) <= ci["# Samples"] or not r[ci["# Samples"]].isdigit():
        continue
    n = int(r[ci["# Samples"]]); tot += n
    src = r[ci["Source"]].strip()
    op = src.split()[1] if src.startswith("@") and len(src.split()) > 1 else src.split()[0]
    anchor = any(k in op for k in ("SYNCS", "LDTM", "STTM", "BRA", "BAR", "UTCHMMA", "UTMALDG", "EXIT", "ATOM", "RED", "CCTL", "MEMBAR", "ERRBAR", "WARPSYNC", "UTCBAR", "ACQBULK"))
    if n >= thr or anchor:
        if accn:
            print("   ... %4d instrs, %5d samples" % (accn, acc))
        acc, accn = 0, 0
        st = sorted(((int(r[ci[c]]), c[6:]) for c in stall_cols if r[ci[c]].isdigit() and int(r[ci[c]]) > 0), reverse=True)[:3]
        print("%6d  x%-7s %-90s %s" % (n, r[ci["Instructions Executed"]], src[:90], " ".join("%s=%d" % (k, v) for v, k in st)))
    else:
        acc += n; accn += 1
if accn:
    print("   ... %4d instrs, %5d samples" % (accn, acc))
print("total", tot)
