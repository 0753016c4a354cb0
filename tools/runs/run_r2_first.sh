#!/bin/bash
# First GPU call of the next round: what was written after round 1's GPU budget was spent — the ColorJitter / GaussianBlur kernels,
# the Trainer with the device input stage, the v2.5 Engine loop — and the first timing of the input stage.  gpurun --timeout 600 -- 'bash tools/runs/run_r2_first.sh'
mkdir -p gpurun_out
python -m pytest tests/test_zz_input_stage_gpu.py tests/test_zzz_engine_gpu.py -q 2>&1 | tail -15 | tee gpurun_out/r02_input_stage_tests.txt
python tools/input_stage_probe.py 2>&1 | tail -12 | tee gpurun_out/r02_input_stage_probe.txt
