#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_kernels_gpu.py tests/test_infonce_tc_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_vit.log 2>&1
echo "== vit + infonce tests rc=$?"; tail -n 6 gpurun_out/r02_test_vit.log
timeout 600 python tools/perf_probe.py vit > gpurun_out/r02_perf_vit_new.log 2>&1; grep "attention" gpurun_out/r02_perf_vit_new.log
for cs in 1 2; do
  PASSL_B200_NCE_CS=$cs timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_probe_cs$cs.log 2>&1
  echo "nce cs=$cs: $(head -c 250 gpurun_out/r02_nce_probe_cs$cs.log)"; sed -n 2,6p gpurun_out/r02_nce_probe_cs$cs.log; grep "loop done\|t3: tile done\|t2: tile done" gpurun_out/r02_nce_probe_cs$cs.log | head -3
done
