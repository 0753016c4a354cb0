#!/bin/bash
# round 2, call 13: the whole GPU suite exactly as the driver runs it, smoke(), input-stage timing, C2 bench line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 12 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02_smoke.log
timeout 300 python tools/input_stage_probe.py > gpurun_out/r02_input_stage_probe.txt 2>&1; echo "input stage rc=$?"; cat gpurun_out/r02_input_stage_probe.txt | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['ms_each_step'], d['roofline_infonce']['frac'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/r02_bench_ref.json
