#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_c.txt 2>&1; echo "probe rc=$?"; grep -E "fc1 fwd|fc2 dgrad|block total|qkv fwd" gpurun_out/r02_vit_gemm_probe_c.txt
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_gemm_epi.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_test_gemm_epi.log
