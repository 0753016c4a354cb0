#!/bin/bash
mkdir -p gpurun_out
PASSL_B200_BENCH_LAUNCH_TABLE=gpurun_out/r02_c2_launch_table.txt timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_tbl.json 2> gpurun_out/r02_bench_c2_tbl.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench_c2_tbl.err; wc -l gpurun_out/r02_c2_launch_table.txt
