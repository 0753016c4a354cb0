#!/bin/bash
mkdir -p gpurun_out
# new instantiation first, under a short limit (a hung kernel must not eat the budget)
timeout 150 python -m pytest tests/test_conv_gpu.py -x -q -m gpu --timeout 100 --timeout-method=thread > gpurun_out/r02_test_conv_first.log 2>&1; rc=$?; echo "conv tests rc=$rc"; tail -3 gpurun_out/r02_test_conv_first.log
if [ $rc -ne 0 ]; then echo "stopping: conv tests failed"; exit 0; fi
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 400 --timeout-method=thread > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 4 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_c2_launch_table_f.txt timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_tbl.json 2> gpurun_out/r02_bench_c2_tbl.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_c2_tbl.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_tbl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
PY
grep "(64, 3, 3, 64)" gpurun_out/r02_c2_launch_table_f.txt | grep -v wgrad | sort -rn | head -8
