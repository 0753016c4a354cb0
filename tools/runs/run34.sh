#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; tail -4 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r01_bench_1gpu.log 2>&1; tail -1 gpurun_out/r01_bench_1gpu.log > gpurun_out/r01_bench_1gpu.json; cut -c1-300 gpurun_out/r01_bench_1gpu.json
