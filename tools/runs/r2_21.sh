#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/r02_test_gemm_epi.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_test_gemm_epi.log
timeout 600 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_f.txt 2>&1; echo "probe rc=$?"; grep -E "fc2 dgrad|fc1 fwd|block total" gpurun_out/r02_vit_gemm_probe_f.txt | head -40
