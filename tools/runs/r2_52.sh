#!/bin/bash
mkdir -p gpurun_out
for c in c4 c5; do
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_${c}_table_ncu.txt timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"gemm_tcgen05|wgrad_halo" --csv --log-file gpurun_out/r02_${c}_gemm_dram.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_${c}_gemm_dram.log 2>&1; echo "ncu $c rc=$?"; wc -l gpurun_out/r02_${c}_gemm_dram.csv gpurun_out/r02_${c}_table_ncu.txt
done
