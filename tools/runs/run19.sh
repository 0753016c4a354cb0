#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_gemm_gpu.py tests/test_conv_gpu.py tests/test_infonce_tc_gpu.py tests/test_resnet_gpu.py 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -30
timeout 300 python tools/perf_probe.py conv 2>&1 | tail -8 | cut -c1-200
timeout 300 python tools/perf_probe.py infonce 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.log').read().strip().splitlines()[-1])
print(d['e2e']); print(d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline_other']['frac'], d['roofline_other']['share_of_step']); print(d['roofline_infonce']['us_per_launch'], d['roofline_infonce']['frac'])
PY
