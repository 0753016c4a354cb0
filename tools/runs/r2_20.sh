#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/r02_test_gemm_epi.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_test_gemm_epi.log
timeout 600 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_e.txt 2>&1; echo "probe rc=$?"; grep -E "fwd|dgrad|block total" gpurun_out/r02_vit_gemm_probe_e.txt | head -40
timeout 300 python tools/bn_probe.py > gpurun_out/r02_bn_probe.txt 2>&1; echo "bn rc=$?"; cat gpurun_out/r02_bn_probe.txt
