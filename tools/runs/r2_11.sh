#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vit_kernels_gpu.py tests/test_mae_gpu.py tests/test_clip_gpu.py -q -m gpu --timeout 300 > gpurun_out/r02_test_vit.log 2>&1
echo "== vit tests (CS auto) rc=$?"; tail -n 8 gpurun_out/r02_test_vit.log
for cs in 1 2; do
  PASSL_B200_ATTN_CS=$cs timeout 600 python tools/perf_probe.py vit > gpurun_out/r02_perf_vit_cs$cs.log 2>&1; echo "CS=$cs"; grep "attention\|MAE" gpurun_out/r02_perf_vit_cs$cs.log
done
timeout 900 python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
echo "bench c3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['ms_each_step'])"
