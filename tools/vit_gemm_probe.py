"""Every GEMM of one ViT block (forward, dgrad, wgrad — with the epilogues the block really uses) at the token counts of the
BASELINE configs, against a plain cuBLAS matmul of the same shape (reference point only).  Developer tool (round 2):
    python tools/vit_gemm_probe.py"""
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


def probe(T, D, Hd, tag):
    """T tokens, model width D, MLP width Hd."""
    x = torch.randn(T, D, device="cuda").bfloat16()
    rows = []

    def one(name, M, N, Kd, fn, ref):
        ms, rs = timeit(fn), timeit(ref)
        fl = 2.0 * M * N * Kd
        rows.append((name, M, N, Kd, ms, fl / ms / 1e9, rs, fl / rs / 1e9))
    for (nm, cin, cout, act, res) in [("qkv", D, 3 * D, None, False), ("proj", D, D, None, True), ("fc1", D, Hd, "gelu", False),
                                      ("fc2", Hd, D, None, True)]:
        a = torch.randn(T, cin, device="cuda").bfloat16()
        w = torch.randn(cout, cin, device="cuda").bfloat16() * 0.02
        bias = torch.zeros(cout, device="cuda")
        r = torch.randn(T, cout, device="cuda").bfloat16() if res else None
        u = torch.empty(T, cout, device="cuda", dtype=torch.bfloat16) if act else None
        out = torch.empty(T, cout, device="cuda", dtype=torch.bfloat16)
        one("%s fwd (+bias%s%s)" % (nm, "+gelu+preact" if act else "", "+residual" if res else ""), T, cout, cin,
            lambda: K.gemm(a, w, bias=bias, act=act, preact_out=u, residual=r, out=out), lambda: torch.matmul(a, w.t(), out=out))
        dy = torch.randn(T, cout, device="cuda").bfloat16()
        dx = torch.empty(T, cin, device="cuda", dtype=torch.bfloat16)
        aux = torch.randn(T, cin, device="cuda").bfloat16() if nm == "fc2" else None
        one("%s dgrad%s" % (nm, " (+gelu')" if aux is not None else ""), T, cin, cout,
            lambda: K.gemm(dy, w, b_t=True, aux=aux, aux_mode_name="gelu_grad", out=dx) if aux is not None else K.gemm(dy, w, b_t=True, out=dx),
            lambda: torch.matmul(dy, w, out=dx))
        dw = torch.zeros(cout, cin, device="cuda")
        dwb = torch.empty(cout, cin, device="cuda", dtype=torch.bfloat16)
        sp = K.wgrad_splits(cout, cin, T)
        one("%s wgrad (split-K %d, fp32 accumulate)" % (nm, sp), cout, cin, T,
            lambda: K.gemm(dy, a, a_t=True, b_t=True, out=dw, accumulate=True, splits=sp), lambda: torch.matmul(dy.t(), a, out=dwb))
    print("== %s: T=%d D=%d mlp=%d" % (tag, T, D, Hd))
    tot, tot_ref = 0.0, 0.0
    for name, M, N, Kd, ms, tf, rs, rtf in rows:
        print("%-44s M%-7d N%-5d K%-7d %8.3f ms %7.0f TF/s | cuBLAS %8.3f ms %7.0f TF/s | %.2fx" % (name, M, N, Kd, ms, tf, rs, rtf, ms / rs))
        tot += ms; tot_ref += rs
    print("block total %.3f ms vs cuBLAS plain matmuls %.3f ms (%.2fx)" % (tot, tot_ref, tot / tot_ref), flush=True)


if __name__ == "__main__":
    probe(512 * 50, 768, 3072, "MAE encoder (C4)")
    probe(512 * 197, 512, 2048, "MAE decoder (C4)")
    probe(1024 * 197, 768, 3072, "CLIP image tower (C5)")
