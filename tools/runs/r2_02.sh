#!/bin/bash
# round 2, call 2: ncu --set full of the single-launch InfoNCE forward / backward (stall sampling per instruction)
mkdir -p gpurun_out
for v in 0 2; do
PASSL_B200_NCE_POLY=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_tc_fwd -s 2 -c 1 -f -o gpurun_out/r02_nce_fwd_poly$v python tools/ncu_target.py infonce > gpurun_out/r02_ncu_nce_fwd$v.log 2>&1
echo "ncu fwd poly$v rc=$?"
done
PASSL_B200_NCE_POLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_tc_bwd -s 2 -c 1 -f -o gpurun_out/r02_nce_bwd python tools/ncu_target.py infonce > gpurun_out/r02_ncu_nce_bwd.log 2>&1
echo "ncu bwd rc=$?"
ls -la gpurun_out/*.ncu-rep
