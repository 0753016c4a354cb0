#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 400 --timeout-method=thread > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 3 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_smoke.log
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_c2_launch_table_h.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2_1gpu.json 2> gpurun_out/r02_bench_c2_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
r=d['roofline']; o=d.get('roofline_other',{})
print('   ', r['bound'], round(r['frac'],3), round(r['share_of_step'],3), '| other', o.get('bound'), round(o.get('frac',0),3), round(o.get('share_of_step',0),3), '| all tflops', round(r.get('all_tflops',0),1))
PY
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; echo "bench c3 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c3_1gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks'])"
