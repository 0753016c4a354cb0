#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_gemm_gpu.py tests/test_conv_gpu.py tests/test_infonce_tc_gpu.py
timeout 600 python tools/perf_probe.py all > gpurun_out/perf_probe.log 2>&1; cat gpurun_out/perf_probe.log | tail -24
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infonce.csv python tools/ncu_target.py infonce > gpurun_out/ncu_infonce.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_infonce.csv | head -4
