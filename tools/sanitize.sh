#!/bin/bash
# compute-sanitizer passes over the kernel tests (the reference has no race / memory checking; these are ours).
# Slow (10-50x): run on a box with time to spare, one tool at a time.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      (or racecheck / synccheck / initcheck)
TOOL=${1:-memcheck}
mkdir -p gpurun_out
# small-shape tests of every kernel family; the big-shape / full-model tests are left out on purpose (hours under the sanitizer)
TESTS="tests/test_gemm_gpu.py tests/test_conv_gpu.py tests/test_simce_gpu.py tests/test_infonce_tc_gpu.py tests/test_vit_kernels_gpu.py \
tests/test_mae_gpu.py tests/test_clip_gpu.py tests/test_optim_gpu.py tests/test_zz_input_stage_gpu.py"
EXISTING=""
for t in $TESTS; do [ -f "$t" ] && EXISTING="$EXISTING $t"; done
timeout 1400 compute-sanitizer --tool "$TOOL" --error-exitcode 1 --target-processes all \
  python -m pytest $EXISTING -x -q -m gpu -k "not bench_shape and not vit_base and not r50" \
  > gpurun_out/sanitizer_${TOOL}.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Error" gpurun_out/sanitizer_${TOOL}.log | tail -8
