#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_d.txt 2>&1; echo "probe rc=$?"; grep -E "fwd|fc2 dgrad|block total" gpurun_out/r02_vit_gemm_probe_d.txt | head -40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 3 -f -o gpurun_out/r02_vitgemm python tools/ncu_target.py vitgemm > gpurun_out/r02_vitgemm_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_vitgemm_ncu.log
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py tests/test_conv_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_gemm_epi.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_test_gemm_epi.log
