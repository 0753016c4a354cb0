#!/bin/bash
# Round-end evidence run (1 GPU): all GPU tests, smoke, micro-benchmarks, bench (+reference arm), ncu launch list and full captures.
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -60
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python tools/perf_probe.py all > gpurun_out/r01_perf_probe.log 2>&1; tail -24 gpurun_out/r01_perf_probe.log | cut -c1-200
timeout 600 python tools/perf_probe.py vit > gpurun_out/r01_perf_vit.log 2>&1; tail -6 gpurun_out/r01_perf_vit.log | cut -c1-200
timeout 300 python tools/nce_timeline.py > gpurun_out/r01_nce_timeline.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r01_bench_1gpu.log 2>&1; tail -1 gpurun_out/r01_bench_1gpu.log > gpurun_out/r01_bench_1gpu.json; cut -c1-400 gpurun_out/r01_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 1200 --csv \
  --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r01_bench_launches.csv --marker stem_pack_input > gpurun_out/r01_bench_launch_summary.txt; head -30 gpurun_out/r01_bench_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 6 -o gpurun_out/r01_c3_full \
  python tools/ncu_target.py c3 > gpurun_out/ncu_c3_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wgrad_halo|bn_bwd_apply|bn_reduce" -c 8 -o gpurun_out/r01_resnet_misc_full \
  python tools/ncu_target.py resnet 64 > gpurun_out/ncu_resnet_misc_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
