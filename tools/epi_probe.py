"""Perf experiment: K-small write-heavy GEMM (layer1 conv3 shape) with parts of the epilogue disabled (PASSL_B200_EPI_DEBUG)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200 import kernels as K
M = 1024 * 56 * 56
x = torch.randn(M, 64, device="cuda").bfloat16()
w = torch.randn(256, 64, device="cuda").bfloat16()
y = torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
part = K.stats_buffer(256, "cuda")
res = torch.randn(M, 256, device="cuda").bfloat16()
w2 = torch.randn(64, 256, device="cuda").bfloat16()
y2 = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
print("dbg=%s  c3 plain %.0f us | +stats %.0f us | +residual %.0f us | c1 (256->64) %.0f us   [floors: 316 / 316 / 569 / 316 us]" % (
    os.environ.get("PASSL_B200_EPI_DEBUG", "0"), t(lambda: K.gemm(x, w, out=y)), t(lambda: K.gemm(x, w, out=y, col_stats=part)),
    t(lambda: K.gemm(x, w, out=y, residual=res)), t(lambda: K.gemm(y, w2, out=y2))))
