#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_infonce_tc_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/r02_test_infonce.log 2>&1
echo "tests rc=$?"; tail -n 5 gpurun_out/r02_test_infonce.log
for v in 0 2; do
  PASSL_B200_NCE_POLY=$v timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_probe_poly$v.log 2>&1
  echo "poly $v rc=$?"; cat gpurun_out/r02_nce_probe_poly$v.log
done
