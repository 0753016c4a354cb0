#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; tail -5 gpurun_out/pytest_gpu_all.log
timeout 600 python -m pytest tests/ -x -q -m "not gpu" 2>&1 | tail -2
