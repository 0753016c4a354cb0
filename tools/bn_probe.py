"""BatchNorm passes of ResNet-50 (per-GPU batch 128 x 2 views = 256 images) against the HBM roofline: algorithmic bytes / time.
Developer tool (round 2):  python tools/bn_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
print("images %d; bytes = algorithmic (each tensor once)" % B)
for (hw, C, res) in [(56, 64, False), (56, 256, True), (28, 128, False), (28, 512, True), (14, 256, False), (14, 1024, True),
                     (7, 512, False), (7, 2048, True)]:
    P = B * hw * hw
    y = torch.randn(P, C, device="cuda").bfloat16()
    r = torch.randn(P, C, device="cuda").bfloat16() if res else None
    dz = torch.randn(P, C, device="cuda").bfloat16()
    gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    part = K.bn_stats(y)
    msss = K.bn_finalize(part, gamma, beta, None, None, P)
    nb = P * C * 2
    t_stats = timeit(lambda: K.bn_stats(y))
    t_apply = timeit(lambda: K.bn_apply_mask(y, msss, residual=r) if res else K.bn_apply(y, msss, True))
    if res:
        z, mask = K.bn_apply_mask(y, msss, residual=r)
        t_bwd = timeit(lambda: K.bn_bwd(y, dz, None, msss, gamma, True, want_dres=True, mask_bits=mask))
        bwd_bytes = 2 * 2 * nb + 2 * nb + 2 * nb // 16     # reduce: y, dz (+mask); apply: y, dz -> dy, dres (+mask)
        app_bytes = 3 * nb + nb // 16
    else:
        t_bwd = timeit(lambda: K.bn_bwd(y, dz, None, msss, gamma, True))
        bwd_bytes = 2 * 2 * nb + nb
        app_bytes = 2 * nb
    print("[%5d x %4d ch%s] stats %7.1f us %5.0f GB/s | apply %7.1f us %5.0f GB/s | bwd (reduce + finalize + apply) %7.1f us %5.0f GB/s" % (
        P, C, " +res" if res else "     ", t_stats * 1e3, nb / t_stats / 1e6, t_apply * 1e3, app_bytes / t_apply / 1e6,
        t_bwd * 1e3, bwd_bytes / t_bwd / 1e6))
