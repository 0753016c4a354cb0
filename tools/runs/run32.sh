#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_resnet_gpu.py tests/test_models_gpu.py tests/test_simclr_gpu.py 2>&1 | grep -E "^==|passed|failed|^E  |Error" | head -30
for i in 1 2; do timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), round(d['value'],1), round(d['e2e']['value'],1), d['clocks'])"; done
