#!/bin/bash
mkdir -p gpurun_out
PASSL_B200_GEMM_HEAVY_PAIR=1 timeout 200 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py -x -q -m gpu --timeout 100 --timeout-method=thread > gpurun_out/r02_test_ew16.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/r02_test_ew16.log
if [ $rc -ne 0 ]; then exit 0; fi
PASSL_B200_GEMM_HEAVY_PAIR=1 timeout 300 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_l.txt 2>&1; echo "probe (heavy pair) rc=$?"; grep -E "fc1 fwd|fc2 dgrad|block total" gpurun_out/r02_vit_gemm_probe_l.txt
timeout 300 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_m.txt 2>&1; echo "probe (default) rc=$?"; grep -E "fc1 fwd|fc2 dgrad|block total" gpurun_out/r02_vit_gemm_probe_m.txt
