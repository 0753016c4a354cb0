#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do PASSL_B200_NO_RES_MMA=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NO_RES_MMA=$v', round(d['ms_per_step'],2), round(d['value'],1), d['clocks'])"; done
