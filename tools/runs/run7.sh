#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_mae_gpu.py tests/test_resnet_gpu.py tests/test_infonce_tc_gpu.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infonce.csv python tools/ncu_target.py infonce > gpurun_out/ncu_infonce.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_infonce.csv | head -5
