#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_vit_kernels_gpu.py tests/test_infonce_tc_gpu.py tests/test_resnet_gpu.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infonce.csv python tools/ncu_target.py infonce > gpurun_out/ncu_infonce.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_infonce.csv | head -6
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-2500
python tools/summarize_launches.py gpurun_out/launches_resnet.csv 2>/dev/null | head -3
