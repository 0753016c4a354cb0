#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_conv_gpu.py -x -q -m gpu --timeout 100 --timeout-method=thread > gpurun_out/r02_test_halo_tma.log 2>&1; rc=$?; echo "conv tests rc=$rc"; tail -5 gpurun_out/r02_test_halo_tma.log
if [ $rc -ne 0 ]; then exit 0; fi
python - <<'PY'
import torch, sys
sys.path.insert(0,'.')
from passl_b200 import kernels as K
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/iters
for (hw,C) in [(56,64),(14,256)]:
    x=torch.randn(1024,hw,hw,C,device='cuda').bfloat16(); w=(torch.randn(C,3,3,C,device='cuda')/(3*C**0.5)).bfloat16()
    y=K.conv2d_fwd(x,w,stride=1,pad=1); dy=torch.randn_like(y)
    part=K.stats_buffer(C,'cuda')
    print("conv3x3 %dch %dx%d B=1024: fwd %.1f us, fwd+stats %.1f us, dgrad %.1f us"%(C,hw,hw,timeit(lambda: K.conv2d_fwd(x,w,stride=1,pad=1,out=y))*1e3, timeit(lambda: K.conv2d_fwd(x,w,stride=1,pad=1,out=y,col_stats=part))*1e3, timeit(lambda: K.conv2d_dgrad(dy,w,tuple(x.shape),stride=1,pad=1))*1e3))
PY
