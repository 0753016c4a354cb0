#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py tests/test_mae_gpu.py tests/test_clip_gpu.py -x -q -m gpu --timeout 100 --timeout-method=thread > gpurun_out/r02_test_ew16.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/r02_test_ew16.log
if [ $rc -ne 0 ]; then exit 0; fi
for c in c4 c5; do
timeout 600 python bench.py --config $c --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${c}_x.json 2> gpurun_out/r02_bench_${c}_x.err; echo "bench rc=$?"
python - gpurun_out/r02_bench_${c}_x.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('  ', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), d['clocks'])
r=d['roofline']; print('   ', r['bound'], round(r['frac'],3), round(r['share_of_step'],3), 'all tflops', round(r.get('all_tflops',0),1))
PY
done
