#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "^==|passed|failed|Error|error" | head -40
timeout 600 python tools/perf_probe.py all > gpurun_out/perf_probe.log 2>&1; tail -45 gpurun_out/perf_probe.log | cut -c1-220
timeout 300 python tools/nce_timeline.py > gpurun_out/nce_timeline.log 2>&1; tail -30 gpurun_out/nce_timeline.log | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | cut -c1-1500
