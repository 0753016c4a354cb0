#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 3 -f -o gpurun_out/r02_vitgemm python tools/ncu_target.py vitgemm > gpurun_out/r02_vitgemm_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_vitgemm_ncu.log
