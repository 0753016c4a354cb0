"""Fused attention forward / backward throughput at the shapes of the BASELINE configs.  Developer tool:  python tools/attn_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200 import kernels_vit as V  # noqa: E402


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


for (B, N, H, d, causal, tag) in [(512, 197, 12, 64, False, "ViT-B/16"), (512, 50, 12, 64, False, "MAE encoder"),
                                  (256, 197, 16, 32, False, "MAE decoder"), (512, 77, 8, 64, True, "CLIP text (causal)")]:
    qkv = torch.randn(B * N, 3 * H * d, device="cuda").bfloat16()
    out, lse = V.attention_fwd(qkv, B, N, H, d, causal=causal)
    dout = torch.randn_like(out)
    f = timeit(lambda: V.attention_fwd(qkv, B, N, H, d, causal=causal))
    b = timeit(lambda: V.attention_bwd(qkv, dout, out, lse, B, N, H, d, causal=causal))
    fl = 4.0 * B * H * N * N * d
    print("%-20s B%d N%d H%d d%d: fwd %.3f ms %5.0f TF/s | bwd %.3f ms %5.0f TF/s (2.5x the forward FLOPs)" %
          (tag, B, N, H, d, f, fl / f / 1e9, b, 2.5 * fl / b / 1e9), flush=True)
