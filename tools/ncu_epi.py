"""Stall samples of one launch of an .ncu-rep restricted to the instructions executed N times (e.g. the epilogue chunk loop):
address-ordered, low-sample runs merged.  usage: ncu_epi.py file.ncu-rep launch_index exec_count[,exec_count...] [min_samples]"""
import csv, io, subprocess, sys
rep, launch, counts = sys.argv[1], sys.argv[2], set(sys.argv[3].split(","))
thr = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
seen, data = set(), []
for r in rows[2:]:
    if len(r) <= ci["# Samples"] or not r[ci["# Samples"]].isdigit() or r[ci["Address"]] in seen:
        continue
    seen.add(r[ci["Address"]]); data.append(r)
tot_all = sum(int(r[ci["# Samples"]]) for r in data)
sel = [r for r in data if r[ci["Instructions Executed"]] in counts]
tot = sum(int(r[ci["# Samples"]]) for r in sel)
agg = {}
for r in sel:
    for c in stall:
        if r[ci[c]].isdigit():
            agg[c[6:]] = agg.get(c[6:], 0) + int(r[ci[c]])
print("selected %d instrs, %d of %d samples; by reason: %s" % (len(sel), tot, tot_all, " ".join("%s=%d" % kv for kv in sorted(agg.items(), key=lambda x: -x[1])[:8])))
acc = cnt = 0
for r in sel:
    n = int(r[ci["# Samples"]]); src = r[ci["Source"]].strip()
    if n >= thr or any(k in src for k in ("LDTM", "WARPSYNC", "SYNCS")):
        if cnt:
            print("      ... %3d instrs %5d samples" % (cnt, acc)); acc = cnt = 0
        st = sorted(((int(r[ci[c]]), c[6:]) for c in stall if r[ci[c]].isdigit() and int(r[ci[c]]) > 0), reverse=True)[:2]
        print("%s %6s %5d  %-64s %s" % (r[ci["Address"]][-5:], r[ci["Instructions Executed"]], n, src[:64], " ".join("%s=%d" % (k, v) for v, k in st)))
    else:
        acc += n; cnt += 1
if cnt:
    print("      ... %3d instrs %5d samples" % (cnt, acc))
