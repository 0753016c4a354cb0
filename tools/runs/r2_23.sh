#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py -q -m gpu --timeout 120 -x > gpurun_out/r02_test_gemm_pair.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_test_gemm_pair.log
timeout 300 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_g.txt 2>&1; echo "probe rc=$?"; grep -E "fwd|dgrad|block total" gpurun_out/r02_vit_gemm_probe_g.txt | head -40; tail -3 gpurun_out/r02_vit_gemm_probe_g.txt
