#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 4 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
