#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_conv_gpu.py tests/test_resnet_gpu.py 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -30
timeout 300 python tools/perf_probe.py conv 2>&1 | tail -8 | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-300
PASSL_B200_NO_HALO=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
