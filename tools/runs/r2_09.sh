#!/bin/bash
# 2-GPU box: peer-memory tests (spawned by the single-process entry), then the 2-GPU bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_peer_gpu.py -q -m gpu --timeout 800 -s > gpurun_out/r02_test_peer.log 2>&1
echo "== peer tests rc=$?"; tail -n 30 gpurun_out/r02_test_peer.log; cat gpurun_out/r02_peer_infonce_latency.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_c2_2gpu.json 2> gpurun_out/r02_bench_c2_2gpu.err
echo "bench 2gpu rc=$?"; head -c 600 gpurun_out/r02_bench_c2_2gpu.json; tail -3 gpurun_out/r02_bench_c2_2gpu.err
