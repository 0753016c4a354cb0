#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_resnet_gpu.py tests/test_simclr_gpu.py tests/test_models_gpu.py 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -30
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-700
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 6 -o gpurun_out/r01_c3_full \
  python tools/ncu_target.py c3 > gpurun_out/ncu_c3_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
