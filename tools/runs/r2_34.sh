#!/bin/bash
# round 2, call 34: evidence run (1 GPU) — whole GPU suite, smoke, probes, ncu captures of the changed kernels, bench lines of all configs
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_pytest_gpu_all.log 2>&1
echo "== pytest -m gpu rc=$?"; tail -n 3 gpurun_out/r02_pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02_smoke.log
timeout 300 python tools/vit_gemm_probe.py > gpurun_out/r02_vit_gemm_probe_after.txt 2>&1; echo "gemm probe rc=$?"; grep "block total" gpurun_out/r02_vit_gemm_probe_after.txt
timeout 300 python tools/attn_probe.py > gpurun_out/r02_attn_probe.txt 2>&1; echo "attn probe rc=$?"
timeout 300 python tools/nce_probe.py timeline > gpurun_out/r02_nce_timeline.txt 2>&1; echo "nce timeline rc=$?"; tail -4 gpurun_out/r02_nce_timeline.txt
for c in c2 c3 c4 c5; do
  extra=""; [ $c != c2 ] && extra="--no-cpu-baseline"
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 $extra > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err; echo "bench $c rc=$?"
  python - gpurun_out/r02_bench_${c}_1gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('  ', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), d['clocks'])
r=d['roofline']; o=d.get('roofline_other',{})
print('   ', r['bound'], round(r['frac'],3), round(r['share_of_step'],3), '| other', o.get('bound'), round(o.get('frac',0),3), round(o.get('share_of_step',0),3), '| all tflops', round(r.get('all_tflops',0),1), '| infonce', round(d['roofline_infonce']['frac'],3))
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; head -c 400 gpurun_out/r02_bench_reference_arm.json; echo
# ncu: launch list of one C2 step, full captures of the ViT GEMMs (pair + epilogue variants) and the attention kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/r02_bench_launches.csv --marker stem_pack_input > gpurun_out/r02_bench_launch_summary.txt 2>&1; head -14 gpurun_out/r02_bench_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 3 -f -o gpurun_out/r02_vitgemm python tools/ncu_target.py vitgemm > gpurun_out/r02_vitgemm_ncu.log 2>&1; echo "ncu vitgemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 8 -f -o gpurun_out/r02_attn python tools/ncu_target.py attn > gpurun_out/r02_attn_ncu.log 2>&1; echo "ncu attn rc=$?"
