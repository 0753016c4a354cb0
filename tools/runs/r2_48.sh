#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_conv_gpu.py -x -q -m gpu --timeout 100 --timeout-method=thread > gpurun_out/r02_test_ew16.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/r02_test_ew16.log
if [ $rc -ne 0 ]; then exit 0; fi
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_c2_launch_table_g.txt timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_tbl.json 2> gpurun_out/r02_bench_c2_tbl.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_c2_tbl.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_tbl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
PY
grep "col_stats" gpurun_out/r02_c2_launch_table_g.txt | grep "1, 1," | sort -rn | head -12
