#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_resnet_gpu.py tests/test_simclr_gpu.py -q -m gpu --timeout 200 > gpurun_out/r02_test_pool.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_test_pool.log
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
N=1024
y=torch.randn(N,112,112,64,device='cuda').bfloat16()
part=K.bn_stats(y.view(-1,64)); msss=K.bn_finalize(part, torch.ones(64,device='cuda'), torch.zeros(64,device='cuda'), None, None, y.numel()//64)
out,arg=K.bn_relu_maxpool_fwd(y,msss)
t=timeit(lambda: K.bn_relu_maxpool_fwd(y,msss)); b=y.numel()*2+out.numel()*3
print("bn_relu_maxpool fwd %.1f us %.0f GB/s"%(t*1e3,b/t/1e6))
dy=torch.randn_like(out)
t=timeit(lambda: K.maxpool_bwd(dy,arg,tuple(y.shape))); b=y.numel()*2+out.numel()*3
print("maxpool bwd %.1f us %.0f GB/s"%(t*1e3,b/t/1e6))
PY
