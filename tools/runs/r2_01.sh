#!/bin/bash
# round 2, call 1: new single-launch InfoNCE forward + tcgen05 backward: tests, timings per exponent mix, timeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_infonce_tc_gpu.py tests/test_simce_gpu.py -q -m gpu --timeout 300 -x > gpurun_out/r02_test_infonce.log 2>&1
echo "tests rc=$?"; tail -n 30 gpurun_out/r02_test_infonce.log
for v in 0 1 2 3; do
  PASSL_B200_NCE_POLY=$v timeout 300 python tools/nce_probe.py $( [ $v = 2 ] && echo timeline ) > gpurun_out/r02_nce_probe_poly$v.log 2>&1
  echo "poly $v rc=$?"; tail -n 20 gpurun_out/r02_nce_probe_poly$v.log
done
