#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/probe_models.py > gpurun_out/r01_probe_models.log 2>&1; cat gpurun_out/r01_probe_models.log | cut -c1-200
