#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py tests/test_mae_gpu.py tests/test_clip_gpu.py -q -m gpu --timeout 200 > gpurun_out/r02_test_gemm_pair.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_test_gemm_pair.log
timeout 300 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_i.txt 2>&1; echo "probe rc=$?"; grep -E "wgrad|block total" gpurun_out/r02_vit_gemm_probe_i.txt
summ() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
r=d['roofline']; o=d.get('roofline_other',{})
print('   ', r['bound'], round(r['frac'],3), round(r['share_of_step'],3), '| other', o.get('bound'), round(o.get('frac',0),3), round(o.get('share_of_step',0),3), '| all tflops', round(r.get('all_tflops',0),1))
PY
}
timeout 600 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c5_pair.json 2> gpurun_out/r02_bench_c5_pair.err; echo "bench rc=$?"; summ gpurun_out/r02_bench_c5_pair.json
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c4_pair.json 2> gpurun_out/r02_bench_c4_pair.err; echo "bench rc=$?"; summ gpurun_out/r02_bench_c4_pair.json
