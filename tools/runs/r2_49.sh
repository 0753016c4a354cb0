#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_2gpu_stdout.json 2> gpurun_out/r02_bench_c2_2gpu.err; echo "bench x2 rc=$?"; wc -l gpurun_out/r02_bench_c2_2gpu_stdout.json; head -c 200 gpurun_out/r02_bench_c2_2gpu_stdout.json; echo; grep -c "NCCL version" gpurun_out/r02_bench_c2_2gpu.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2_stdout.json 2> gpurun_out/r02_bench_c2_stdout.err; echo "bench x1 rc=$?"; wc -l gpurun_out/r02_bench_c2_stdout.json; head -c 120 gpurun_out/r02_bench_c2_stdout.json; echo
