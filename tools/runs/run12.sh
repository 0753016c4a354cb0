#!/bin/bash
# ncu launch list of one bench step (share-of-step evidence) + one --set full capture of the conv target (traffic)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 6000 --launch-count 2820 --csv \
  --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_bench_launches.csv | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 6 -o gpurun_out/r01_conv_full \
  python tools/ncu_target.py conv > gpurun_out/ncu_conv_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
