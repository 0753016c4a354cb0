#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 4 -o gpurun_out/r01_c3res_full \
  python tools/ncu_target.py c3res > gpurun_out/ncu_c3res_full.log 2>&1
ls -la gpurun_out/r01_c3res_full.ncu-rep
