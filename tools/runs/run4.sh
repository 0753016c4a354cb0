#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_resnet_gpu.py tests/test_simclr_gpu.py tests/test_simce_gpu.py tests/test_infonce_tc_gpu.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -4 gpurun_out/bench.log
timeout 600 python tools/perf_probe.py resnet > gpurun_out/perf_resnet.log 2>&1; tail -3 gpurun_out/perf_resnet.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_resnet.csv python tools/ncu_target.py resnet 256 > gpurun_out/ncu_resnet.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_resnet.csv | head -32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_tc -c 1 -o gpurun_out/prof_infonce python tools/ncu_target.py infonce > gpurun_out/ncu_infonce_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_infonce.csv python tools/ncu_target.py infonce > gpurun_out/ncu_infonce.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_infonce.csv | head -6
ls -la gpurun_out | head -40
