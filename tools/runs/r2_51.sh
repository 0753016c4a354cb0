#!/bin/bash
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 1 --target-processes all python -m pytest tests/test_gemm_gpu.py -x -q -m gpu -k "sixteen or cta_pair or linear_epilogue" > gpurun_out/r02_sanitizer_memcheck_gemm.log 2>&1; echo "memcheck gemm rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r02_sanitizer_memcheck_gemm.log | tail -6
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 1 --target-processes all python -m pytest tests/test_conv_gpu.py -x -q -m gpu -k "test_conv_fwd or test_conv_dgrad" > gpurun_out/r02_sanitizer_memcheck_conv.log 2>&1; echo "memcheck conv rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r02_sanitizer_memcheck_conv.log | tail -6
