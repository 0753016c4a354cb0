#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_resnet_gpu.py tests/test_trajectory_gpu.py tests/test_zzz_engine_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -s > gpurun_out/r02_$b.log 2>&1
  echo "== $b rc=$?"; tail -n 8 gpurun_out/r02_$b.log
done
cat gpurun_out/r02_resnet50_ResNet_unit_parity.txt; tail -3 gpurun_out/r02_resnet50_ResNet_grad_report.txt; grep trajectory gpurun_out/r02_test_trajectory_gpu.log | head -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02_bench_c2.json; tail -5 gpurun_out/r02_bench_c2.err
