#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02_vit_gemm_probe.txt
timeout 600 python -m pytest tests/test_resnet_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_resnet_gpu.log 2>&1; echo "resnet tests rc=$?"; tail -3 gpurun_out/r02_test_resnet_gpu.log
