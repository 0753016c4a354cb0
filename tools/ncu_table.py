"""One line per profiled launch of an .ncu-rep: duration, tensor-pipe %, issue-active %, DRAM %, DRAM bytes, registers.
usage: ncu_table.py file.ncu-rep [label ...]   (labels are attached to the launches in order)"""
import csv, io, subprocess, sys
rep = sys.argv[1]
labels = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
def g(r, k, d="-"):
    return r[ci[k]] if k in ci else d
print("%-34s %-26s %10s %9s %9s %8s %12s %12s %5s" % ("launch", "kernel", "us", "tensor%", "issue%", "dram%", "dram rd MB", "dram wr MB", "regs"))
for n, r in enumerate(rows[2:]):
    name = g(r, "Kernel Name")[:26]
    dur = float(g(r, "gpu__time_duration.sum", "0").replace(",", ""))
    unit = rows[1][ci["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in ci else "us"
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    def b(k):
        v = float(g(r, k, "0").replace(",", ""))
        u = rows[1][ci[k]] if k in ci else "byte"
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    print("%-34s %-26s %10.1f %9s %9s %8s %12.2f %12.2f %5s" % (
        labels[n] if n < len(labels) else "#%d" % n, name, dur_us,
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")[:6],
        g(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed")[:6],
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[:6],
        b("dram__bytes_read.sum"), b("dram__bytes_write.sum"), g(r, "launch__registers_per_thread")))
