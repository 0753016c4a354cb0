#!/bin/bash
# First GPU call of the next round: the input-stage kernels added after round 1's GPU budget was spent (ColorJitter, GaussianBlur,
# Trainer with the device input stage) and their first timing.  gpurun --timeout 600 -- 'bash tools/runs/run_r2_first.sh'
mkdir -p gpurun_out
python -m pytest tests/test_zz_input_stage_gpu.py -q 2>&1 | tail -15 | tee gpurun_out/r02_input_stage_tests.txt
python tools/input_stage_probe.py 2>&1 | tail -12 | tee gpurun_out/r02_input_stage_probe.txt
