#!/bin/bash
# Run every GPU test file in its own process (a device trap must not poison the others); logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in ${@:-tests/test_*_gpu.py}; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -q -m gpu --timeout 300 > gpurun_out/$b.log 2>&1
  r=$?
  echo "== $b rc=$r"; tail -n 25 gpurun_out/$b.log
  [ $r -ne 0 ] && rc=$r
done
exit $rc
