#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_vit_kernels_gpu.py tests/test_conv_gpu.py -q -m gpu --timeout 200 > gpurun_out/r02_test_gemm_pair.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r02_test_gemm_pair.log
timeout 300 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_g.txt 2>&1; echo "probe rc=$?"; grep -E "block total" gpurun_out/r02_vit_gemm_probe_g.txt
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r02_bench_c2_pair.json 2> gpurun_out/r02_bench_c2_pair.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_c2_pair.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','clocks')}); print(d['roofline']['frac'], d['roofline']['all_tflops'], d.get('roofline_other',{}).get('achieved'))
PY
