#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/pipes > gpurun_out/r02_pipes.txt 2>&1; cat gpurun_out/r02_pipes.txt
PASSL_B200_NCE_POLY=2 timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_probe_poly2.log 2>&1
echo "probe rc=$?"; head -c 300 gpurun_out/r02_nce_probe_poly2.log; echo; sed -n 2,6p gpurun_out/r02_nce_probe_poly2.log; sed -n 22,34p gpurun_out/r02_nce_probe_poly2.log
timeout 600 python -m pytest tests/test_infonce_tc_gpu.py tests/test_vit_kernels_gpu.py tests/test_mae_gpu.py tests/test_clip_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_infonce.log 2>&1
echo "== tests rc=$?"; tail -n 6 gpurun_out/r02_test_infonce.log
timeout 600 python tools/perf_probe.py vit > gpurun_out/r02_perf_vit_new.log 2>&1; echo "perf vit new rc=$?"; grep "attention\|MAE" gpurun_out/r02_perf_vit_new.log
