#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_infonce_tc_gpu.py -q -m gpu -x --timeout 120 > gpurun_out/test_infonce_tc_gpu.log 2>&1; tail -3 gpurun_out/test_infonce_tc_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:infonce_tc -s 2 -c 1 -o gpurun_out/prof_infonce python tools/ncu_target.py infonce > gpurun_out/ncu_infonce_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 3 -o gpurun_out/prof_conv python tools/ncu_target.py conv > gpurun_out/ncu_conv_full.log 2>&1
timeout 600 python tools/perf_probe.py resnet > gpurun_out/perf_resnet.log 2>&1; tail -3 gpurun_out/perf_resnet.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-600
