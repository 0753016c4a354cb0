#!/bin/bash
mkdir -p gpurun_out
for c in c4 c5; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err; echo "bench $c rc=$?"
  python - gpurun_out/r02_bench_${c}_1gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('  ', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), d['clocks'])
r=d['roofline']; print('   ', r['bound'], round(r['frac'],3), round(r['share_of_step'],3), 'traffic', r['traffic'], 'all tflops', round(r.get('all_tflops',0),1))
PY
done
