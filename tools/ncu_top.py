"""Print key metrics + top stall instructions from an .ncu-rep (needs ncu on PATH). usage: ncu_top.py file.ncu-rep [kernel_index]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
for r in rows[2:]:
    print("----")
    for h, v in zip(hdr, r):
        if h in want:
            print("  %-80s %s" % (h, v[:90]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for bi, blk in enumerate(blocks[1:]):
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    if len(rows) < 3:
        continue
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > ci['# Samples'] and r[ci['# Samples']].isdigit()]
    tot = sum(int(r[ci['# Samples']]) for r in data)
    print("==== kernel %d: %s | total samples %d" % (bi, rows[0][1][:70], tot))
    for r in sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:14]:
        print("  %5s %9s  %s" % (r[ci['# Samples']], r[ci['Instructions Executed']], r[ci['Source']][:96]))
