#!/bin/bash
# round 2, call 4: InfoNCE prologue fix (timeline), quantisation-matched parity tests, trajectories, optimizer control word, engine accumulation
mkdir -p gpurun_out
PASSL_B200_NCE_POLY=2 timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_probe_poly2.log 2>&1
echo "probe rc=$?"; head -c 700 gpurun_out/r02_nce_probe_poly2.log; echo; sed -n 2,8p gpurun_out/r02_nce_probe_poly2.log
for f in tests/test_infonce_tc_gpu.py tests/test_optim_gpu.py tests/test_resnet_gpu.py tests/test_trajectory_gpu.py tests/test_zzz_engine_gpu.py tests/test_simce_gpu.py tests/test_models_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -s > gpurun_out/r02_$b.log 2>&1
  echo "== $b rc=$?"; tail -n 12 gpurun_out/r02_$b.log
done
