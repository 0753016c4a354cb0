#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_umma_probe_gpu.py tests/test_conv_gpu.py tests/test_infonce_tc_gpu.py tests/test_resnet_gpu.py tests/test_simclr_gpu.py tests/test_models_gpu.py 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -60
cat gpurun_out/umma_probe_report.txt
timeout 300 python tools/perf_probe.py membw 2>&1 | tail -5
timeout 300 python tools/perf_probe.py infonce 2>&1 | tail -3
timeout 300 python tools/nce_timeline.py > gpurun_out/nce_timeline.log 2>&1; tail -16 gpurun_out/nce_timeline.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-400
