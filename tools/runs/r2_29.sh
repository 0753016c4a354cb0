#!/bin/bash
# round 2, call 29: whole GPU suite + smoke after the GEMM epilogue / CTA-pair work; kernel shares of one C5 and one C4 step
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 6 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02_smoke.log
for c in c5 c4; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_${c}_launches.csv python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_${c}_under_ncu.log 2>&1; echo "ncu $c rc=$?"
python tools/summarize_launches.py gpurun_out/r02_bench_${c}_launches.csv > gpurun_out/r02_bench_${c}_launch_summary.txt 2>&1; head -30 gpurun_out/r02_bench_${c}_launch_summary.txt
done
