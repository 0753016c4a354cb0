#!/bin/bash
mkdir -p gpurun_out
c=c5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_${c}_launches.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_${c}_under_ncu.log 2>&1; echo "ncu $c rc=$?"
python tools/summarize_launches.py gpurun_out/r02_bench_${c}_launches.csv > gpurun_out/r02_bench_${c}_launch_summary_end.txt 2>&1; head -24 gpurun_out/r02_bench_${c}_launch_summary_end.txt
