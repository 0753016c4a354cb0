#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu --timeout 200 -x > gpurun_out/r02_test_halo.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02_test_halo.log
