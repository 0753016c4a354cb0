#!/bin/bash
mkdir -p gpurun_out
timeout 60 python tools/perf_probe.py gemm > gpurun_out/r02_perf_probe_gemm.txt 2>&1; echo "gemm rc=$?"; cat gpurun_out/r02_perf_probe_gemm.txt | grep "^gemm"
timeout 60 python tools/perf_probe.py conv > gpurun_out/r02_perf_probe_conv.txt 2>&1; echo "conv rc=$?"; grep "^conv" gpurun_out/r02_perf_probe_conv.txt
