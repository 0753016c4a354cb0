#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_${N}gpu.log > gpurun_out/r01_bench_${N}gpu.json; cut -c1-330 gpurun_out/r01_bench_${N}gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m pytest tests/test_simclr_gpu.py -q -m gpu -k bench_shape > gpurun_out/dist_pytest.log 2>&1; tail -2 gpurun_out/dist_pytest.log
