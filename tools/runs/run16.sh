#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -60
timeout 600 python tools/perf_probe.py all > gpurun_out/perf_probe.log 2>&1; tail -22 gpurun_out/perf_probe.log | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-1900
