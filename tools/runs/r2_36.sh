#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py -q -m gpu --timeout 200 > gpurun_out/r02_test_wgrad.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_test_wgrad.log
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_c2_launch_table_d.txt timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_tbl.json 2> gpurun_out/r02_bench_c2_tbl.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_c2_tbl.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_tbl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks'])
PY
grep "conv2d_wgrad" gpurun_out/r02_c2_launch_table_d.txt | grep "3, 3," | sort -rn | head -16
