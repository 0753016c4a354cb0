#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_gpu.py -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_c2_2gpu.json 2> gpurun_out/r02_bench_c2_2gpu.err; echo "bench x2 rc=$?"; wc -l gpurun_out/r02_bench_c2_2gpu.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_c2_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['clocks'])"
