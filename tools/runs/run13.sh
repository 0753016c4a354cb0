#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_clip_gpu.py 2>&1 | tail -60 | cut -c1-250
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 564 --csv \
  --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_bench_launches.csv | head -45
