#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; tail -1 gpurun_out/bench_2gpu.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m pytest tests/test_simclr_gpu.py -q -m gpu -k bench_shape > gpurun_out/dist_pytest.log 2>&1; tail -3 gpurun_out/dist_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_2gpu.log 2>&1; tail -1 gpurun_out/bench_ref_2gpu.log | cut -c1-500
