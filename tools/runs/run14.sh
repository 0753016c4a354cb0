#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_clip_gpu.py tests/test_resnet_gpu.py tests/test_simclr_gpu.py tests/test_models_gpu.py 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -40
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-2500
PASSL_B200_SIDE_STREAM=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_noside.log 2>&1; tail -1 gpurun_out/bench_1gpu_noside.log | cut -c1-300
