#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_models_gpu.py
timeout 600 python tools/perf_probe.py vit > gpurun_out/perf_vit.log 2>&1; tail -6 gpurun_out/perf_vit.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; tail -2 gpurun_out/bench_2gpu.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m pytest tests/test_simclr_gpu.py -q -m gpu -k bench_shape > gpurun_out/dist_pytest.log 2>&1; tail -3 gpurun_out/dist_pytest.log
