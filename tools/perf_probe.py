"""Quick GPU timing probe (CUDA events, warm-up, L2 flush between iterations). Prints one line per item."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200 import kernels as K  # noqa: E402

FLUSH = None


def timeit(fn, iters=10, warmup=3, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    res = {}
    if what in ("membw",):
        # directional HBM bandwidth (torch library kernels, measurement only): pure read, pure write, copy, on 4 GiB buffers
        n = 1 << 30
        a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
        b = torch.empty_like(a)
        t = timeit(lambda: a.sum(), flush=False)
        print("read  4 GiB: %.3f ms  %.0f GB/s" % (t, 4 * n / t / 1e6))
        t = timeit(lambda: b.zero_(), flush=False)
        print("write 4 GiB: %.3f ms  %.0f GB/s" % (t, 4 * n / t / 1e6))
        t = timeit(lambda: b.copy_(a), flush=False)
        print("copy  4+4 GiB: %.3f ms  %.0f GB/s" % (t, 8 * n / t / 1e6))
        y = torch.randn(1024 * 56 * 56, 256, device="cuda").bfloat16()
        t = timeit(lambda: K.bn_stats(y), flush=False)
        print("bn_stats read 1.64 GB: %.3f ms %.0f GB/s" % (t, y.numel() * 2 / t / 1e6))
    if what in ("all", "gemm"):
        for (M, N, K_) in [(8192, 8192, 8192), (4096, 4096, 4096), (100352, 64, 576), (100352, 256, 64), (25088, 512, 128),
                           (16384, 2304, 768), (16384, 768, 3072), (16384, 3072, 768)]:
            a = torch.randn(M, K_, device="cuda").bfloat16()
            b = torch.randn(N, K_, device="cuda").bfloat16()
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            ms = timeit(lambda: K.gemm(a, b, out=out))
            ref = timeit(lambda: torch.matmul(a, b.t(), out=out))
            res["gemm_%dx%dx%d" % (M, N, K_)] = dict(ms=ms, tflops=2 * M * N * K_ / ms / 1e9, cublas_ms=ref,
                                                     cublas_tflops=2 * M * N * K_ / ref / 1e9)
            print("gemm", M, N, K_, "%.3f ms %.1f TF/s | cuBLAS %.3f ms %.1f TF/s" % (ms, 2 * M * N * K_ / ms / 1e9, ref, 2 * M * N * K_ / ref / 1e9), flush=True)
    if what in ("all", "conv"):
        B = 64
        for (H, Cin, Cout, R, s, p) in [(56, 64, 64, 3, 1, 1), (56, 128, 128, 3, 2, 1), (28, 128, 128, 3, 1, 1), (14, 256, 256, 3, 1, 1),
                                        (7, 512, 512, 3, 1, 1), (56, 256, 64, 1, 1, 0), (14, 1024, 256, 1, 1, 0)]:
            x = torch.randn(B, H, H, Cin, device="cuda").bfloat16()
            w = torch.randn(Cout, R, R, Cin, device="cuda").bfloat16()
            Ho = (H + 2 * p - R) // s + 1
            fl = 2.0 * B * Ho * Ho * Cout * R * R * Cin
            y = K.conv2d_fwd(x, w, stride=s, pad=p)
            dy = torch.randn_like(y)
            cs = K.stats_buffer(Cout, "cuda")
            f = timeit(lambda: K.conv2d_fwd(x, w, stride=s, pad=p, out=y))
            fs = timeit(lambda: K.conv2d_fwd(x, w, stride=s, pad=p, out=y, col_stats=cs))
            dxb = torch.empty_like(x)
            d = timeit(lambda: K.conv2d_dgrad(dy, w, tuple(x.shape), stride=s, pad=p, out=dxb))
            dw = torch.zeros(Cout, R, R, Cin, device="cuda")
            g = timeit(lambda: K.conv2d_wgrad(x, dy, tuple(w.shape), stride=s, pad=p, out=dw, accumulate=True))
            print("conv H%d %d->%d k%d s%d: fwd %.3f ms %.0f TF/s (with stats %.3f) | dgrad %.3f ms %.0f TF/s | wgrad %.3f ms %.0f TF/s" %
                  (H, Cin, Cout, R, s, f, fl / f / 1e9, fs, d, fl / d / 1e9, g, fl / g / 1e9), flush=True)
            res["conv_%d_%d_%d_%d_%d" % (H, Cin, Cout, R, s)] = dict(fwd_ms=f, fwd_stats_ms=fs, dgrad_ms=d, wgrad_ms=g, gflop=fl / 1e9)
    if what in ("all", "infonce"):
        N, D, Kq, T = 256, 128, 65536, 0.2
        q = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
        k = torch.nn.functional.normalize(torch.randn(N, D, device="cuda"), dim=1)
        queue = torch.nn.functional.normalize(torch.randn(Kq, D, device="cuda"), dim=1).bfloat16()
        qb = q.bfloat16()
        ms = timeit(lambda: K.infonce_tc_fwd(qb, queue, pos=k, scale=1 / T), iters=20)
        by = (2 * N * D + D * Kq) * 2 + 4 * N
        print("infonce_tc fwd N%d K%d: %.4f ms  %.0f GB/s algorithmic (incl. finalize kernel)" % (N, Kq, ms, by / ms / 1e6), flush=True)
        out, lse, tgt, _ = K.infonce_tc_fwd(qb, queue, pos=k, scale=1 / T)
        ms2 = timeit(lambda: K.simce_bwd(q, queue, lse, tgt, pos=k, scale=1 / T), iters=5)
        ms3 = timeit(lambda: K.simce_fwd(q, queue, pos=k, scale=1 / T), iters=5)
        print("simce (SIMT) fwd %.3f ms, bwd %.3f ms" % (ms3, ms2), flush=True)
        res["infonce"] = dict(tc_fwd_ms=ms, simt_fwd_ms=ms3, simt_bwd_ms=ms2)
    if what in ("all", "resnet"):
        from passl_b200.modeling import build_backbone, build_neck
        from passl_b200.core import ParamStore
        import torch.nn as nn
        for B in (64, 256):
            net = nn.Sequential(build_backbone(dict(name="ResNet", depth=50)), build_neck(dict(name="NonLinearNeckV1", in_channels=2048, hid_channels=2048, out_channels=128))).cuda()
            ParamStore(net)
            img = torch.randn(B, 3, 224, 224, device="cuda")

            def step():
                net._param_store.zero_grad()
                e = net(img)
                e.backward(torch.ones_like(e))
            ms = timeit(step, iters=5, warmup=2, flush=False)

            def fwd():
                with torch.no_grad():
                    net(img)
            msf = timeit(fwd, iters=5, warmup=2, flush=False)
            print("resnet50 B=%d fwd+bwd %.2f ms -> %.0f img/s ; fwd only %.2f ms -> %.0f img/s ; mem %.1f GB" %
                  (B, ms, B / ms * 1e3, msf, B / msf * 1e3, torch.cuda.max_memory_allocated() / 2**30), flush=True)
            res["resnet50_B%d" % B] = dict(fwdbwd_ms=ms, fwd_ms=msf)
            del net
            torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/perf_probe_%s.json" % what, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "vit"):
    main()


def probe_vit():
    import torch
    from passl_b200 import kernels_vit as V
    res = {}
    for (B, N, H, d) in [(512, 197, 12, 64), (512, 50, 12, 64), (256, 197, 16, 32)]:
        qkv = torch.randn(B * N, 3 * H * d, device="cuda").bfloat16()
        out, lse = V.attention_fwd(qkv, B, N, H, d)
        dout = torch.randn_like(out)
        f = timeit(lambda: V.attention_fwd(qkv, B, N, H, d), iters=10)
        b = timeit(lambda: V.attention_bwd(qkv, dout, out, lse, B, N, H, d), iters=10)
        fl = 4.0 * B * H * N * N * d
        print("attention B%d N%d H%d d%d: fwd %.3f ms %.0f TF/s | bwd %.3f ms %.0f TF/s (2.5x fwd flops)" %
              (B, N, H, d, f, fl / f / 1e9, b, 2.5 * fl / b / 1e9), flush=True)
        res["attn_%d_%d_%d_%d" % (B, N, H, d)] = dict(fwd_ms=f, bwd_ms=b, fwd_tflops=fl / f / 1e9)
    from passl_b200.core import ParamStore
    from passl_b200.models import build_model
    from passl_b200.optimizer import AdamW
    for B in (128, 512):
        m = build_model(dict(name="mae_vit_base_patch16", norm_pix_loss=True)).cuda()
        st = ParamStore(m)
        opt = AdamW(st, lr=1.5e-4, beta2=0.95, weight_decay=0.05, one_dim_no_decay=True)
        imgs = torch.randn(B, 3, 224, 224, device="cuda")

        def step():
            opt.clear_grad()
            loss, _, _ = m(imgs, 0.75)
            loss.backward()
            opt.step()
        ms = timeit(step, iters=5, warmup=2, flush=False)
        print("MAE ViT-B/16 B=%d full step %.2f ms -> %.0f img/s ; mem %.1f GB" % (B, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 2**30), flush=True)
        res["mae_B%d" % B] = dict(step_ms=ms, ips=B / ms * 1e3)
        del m, st, opt
        torch.cuda.empty_cache()
    import json
    json.dump(res, open("gpurun_out/perf_probe_vit.json", "w"), indent=1)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "vit":
    probe_vit()
