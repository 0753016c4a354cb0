#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 -m pytest tests/test_peer_gpu.py -x -q -s -m gpu > gpurun_out/peer_pytest.log 2>&1; tail -25 gpurun_out/peer_pytest.log | cut -c1-220
cat gpurun_out/peer_exchange_latency.txt
