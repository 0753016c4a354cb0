#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_resnet_gpu.py
timeout 300 python tools/diag_resnet.py 16 128 > gpurun_out/diag_resnet.log 2>&1; tail -22 gpurun_out/diag_resnet.log
timeout 600 python tools/perf_probe.py resnet > gpurun_out/perf_resnet.log 2>&1; tail -5 gpurun_out/perf_resnet.log
# per-kernel device time of one ResNet-50 fwd+bwd (B=64) and of the fused InfoNCE
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_resnet.csv python tools/ncu_target.py resnet > gpurun_out/ncu_resnet.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infonce.csv python tools/ncu_target.py infonce > gpurun_out/ncu_infonce.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_resnet.csv | head -40
python tools/summarize_launches.py gpurun_out/launches_infonce.csv | head -12
