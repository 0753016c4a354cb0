"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line (needs -lineinfo + --import-source on).
usage: ncu_lines.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header row with "# Samples"
files = {}
cur, hdr = None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if cur and hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        n = d.get("# Samples", "0")
        if n.isdigit() and int(n) > 0:
            files.setdefault(cur, []).append((int(n), int(r[0]), d.get("Source", "")[:110], d))
tot = sum(n for f in files.values() for n, *_ in f)
print("total samples", tot)
for f, lst in files.items():
    print("==", f, sum(n for n, *_ in lst))
    for n, ln, src, d in sorted(lst, reverse=True)[:top]:
        stalls = sorted(((int(v), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0), reverse=True)[:3]
        print("  %5d  L%-4d %-110s %s" % (n, ln, src.strip(), " ".join("%s=%d" % (k[6:], v) for v, k in stalls)))
