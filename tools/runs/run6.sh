#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_vit_kernels_gpu.py tests/test_infonce_tc_gpu.py tests/test_simce_gpu.py tests/test_resnet_gpu.py
head -30 gpurun_out/resnet_e2e_grad_report.txt; tail -12 gpurun_out/resnet_e2e_grad_report.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infonce.csv python tools/ncu_target.py infonce > gpurun_out/ncu_infonce.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_infonce.csv | head -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_tc -c 1 -o gpurun_out/prof_infonce python tools/ncu_target.py infonce > gpurun_out/ncu_infonce_full.log 2>&1
