#!/bin/bash
# round 2, call 35 (2 GPUs): the multi-GPU tests (peer exchange, peer-sharded InfoNCE, NCCL paths) and the 2-GPU bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_peer_gpu.py tests/test_models_gpu.py -q -m gpu --timeout 600 > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_pytest_2gpu.log
for c in c2 c4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $c --steps 10 --warmup 3 > gpurun_out/r02_bench_${c}_2gpu.json 2> gpurun_out/r02_bench_${c}_2gpu.err; echo "bench $c x2 rc=$?"
python - gpurun_out/r02_bench_${c}_2gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('  ', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
done
