#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 4 -c 1 -f -o gpurun_out/r02_attn_bwd python tools/ncu_target.py attn > gpurun_out/r02_attn_bwd_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_attn_bwd_ncu.log
