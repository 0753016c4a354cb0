#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_kernels_gpu.py tests/test_mae_gpu.py tests/test_clip_gpu.py -q -m gpu --timeout 200 > gpurun_out/r02_test_attn.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_test_attn.log
timeout 300 python tools/attn_probe.py > gpurun_out/r02_attn_probe.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r02_attn_probe.txt
