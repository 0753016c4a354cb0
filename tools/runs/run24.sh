#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 1200 --csv \
  --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_bench_launches.csv --marker stem_pack_input | head -48
