#!/bin/bash
mkdir -p gpurun_out
PASSL_B200_NCE_POLY=2 timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_probe_poly2.log 2>&1
echo "probe rc=$?"; head -c 420 gpurun_out/r02_nce_probe_poly2.log; echo; sed -n 2,9p gpurun_out/r02_nce_probe_poly2.log; grep -n "producer\|mma" gpurun_out/r02_nce_probe_poly2.log | head -8
timeout 600 python -m pytest tests/test_infonce_tc_gpu.py tests/test_zzz_engine_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/r02_test_infonce.log 2>&1
echo "== tests rc=$?"; tail -n 6 gpurun_out/r02_test_infonce.log
timeout 300 python tools/diag_bwd.py 1024 256 1 0 16 8 > gpurun_out/r02_diag_bwd_1024.log 2>&1; cat gpurun_out/r02_diag_bwd_1024.log | tail -12
timeout 300 python tools/diag_bwd.py 256 64 1 0 16 32 > gpurun_out/r02_diag_bwd_256.log 2>&1; cat gpurun_out/r02_diag_bwd_256.log | tail -12
