#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:"gemm_tcgen05|wgrad_halo" --csv --log-file gpurun_out/r02_conv3_dram_halo.csv python tools/ncu_target.py conv3 > gpurun_out/r02_conv3_dram.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/r02_conv3_dram_halo.csv') if l.startswith('"')]
per=collections.OrderedDict()
for r in csv.DictReader(lines):
    per.setdefault(r['ID'],{'k':r['Kernel Name'][:58]})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for i,d in per.items():
    print("%-58s %7.1f us dram rd %7.1f wr %6.1f | L2->SM %7.1f MB (%5.1f TB/s) hit %4.1f"%(d['k'],d['gpu__time_duration.sum']/1e3,d['dram__bytes_read.sum']/1e6,d['dram__bytes_write.sum']/1e6,d['l1tex__m_xbar2l1tex_read_bytes.sum']/1e6,d['l1tex__m_xbar2l1tex_read_bytes.sum']/d['gpu__time_duration.sum']/1e3,d['lts__t_sector_hit_rate.pct']))
PY
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_c2_launch_table_e.txt timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_tbl.json 2> gpurun_out/r02_bench_c2_tbl.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_c2_tbl.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_tbl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
PY
grep "3, 3," gpurun_out/r02_c2_launch_table_e.txt | grep -v wgrad | sort -rn | head -24
