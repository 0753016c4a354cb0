#!/bin/bash
# round 2, call 5: InfoNCE with L2 hints; pipelined attention fwd (tests + timing, old vs new); reworked parity tests
mkdir -p gpurun_out
PASSL_B200_NCE_POLY=2 timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_probe_poly2.log 2>&1
echo "probe rc=$?"; head -c 420 gpurun_out/r02_nce_probe_poly2.log; echo; sed -n 2,9p gpurun_out/r02_nce_probe_poly2.log; grep -n "producer\|mma" gpurun_out/r02_nce_probe_poly2.log | head -8
timeout 600 python -m pytest tests/test_vit_kernels_gpu.py tests/test_mae_gpu.py tests/test_clip_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_vit.log 2>&1
echo "== vit tests rc=$?"; tail -n 8 gpurun_out/r02_test_vit.log
timeout 600 python tools/perf_probe.py vit > gpurun_out/r02_perf_vit_new.log 2>&1; echo "perf vit new rc=$?"; grep attention gpurun_out/r02_perf_vit_new.log
PASSL_B200_ATTN_V1=1 timeout 600 python tools/perf_probe.py vit > gpurun_out/r02_perf_vit_v1.log 2>&1; grep attention gpurun_out/r02_perf_vit_v1.log
for f in tests/test_resnet_gpu.py tests/test_trajectory_gpu.py tests/test_zzz_engine_gpu.py tests/test_models_gpu.py tests/test_optim_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -s > gpurun_out/r02_$b.log 2>&1
  echo "== $b rc=$?"; tail -n 12 gpurun_out/r02_$b.log
done
cat gpurun_out/r02_resnet50_ResNet_unit_parity.txt
