#!/bin/bash
# round 2, call 10: bench lines of the other BASELINE configs, InfoNCE exponent-mix A/B, ncu captures (attention, InfoNCE), launch list of the c2 step
mkdir -p gpurun_out
for c in c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  echo "bench $c rc=$?"; head -c 330 gpurun_out/r02_bench_$c.json; echo; tail -2 gpurun_out/r02_bench_$c.err
done
for v in 0 1 2 3; do
  PASSL_B200_NCE_POLY=$v timeout 300 python tools/nce_probe.py > gpurun_out/r02_nce_probe_poly$v.log 2>&1
  echo "poly $v: $(head -c 260 gpurun_out/r02_nce_probe_poly$v.log)"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 8 -f -o gpurun_out/r02_attn python tools/ncu_target.py attn > gpurun_out/r02_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_tc_fwd -s 2 -c 1 -f -o gpurun_out/r02_nce_fwd python tools/ncu_target.py infonce > gpurun_out/r02_ncu_nce_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:infonce_tc_bwd -s 2 -c 1 -f -o gpurun_out/r02_nce_bwd python tools/ncu_target.py infonce > gpurun_out/r02_ncu_nce_bwd.log 2>&1
echo "ncu nce rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1300 --launch-count 1300 --csv \
  --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
