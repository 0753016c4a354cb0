#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu --timeout 100 --timeout-method=thread > gpurun_out/r02_test_ew16.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -8 gpurun_out/r02_test_ew16.log
