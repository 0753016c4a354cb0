"""Full training-step throughput of the other BASELINE configurations on one B200 (synthetic data, per-GPU batch of the config):
C3 MoCo v2 ResNet-50 (bs 256, K=65536, Momentum), C5 CLIP ViT-B/16 (image-text pairs, AdamW), MoCo v3 ViT-B/16 (bs 256, AdamW).
C2 (SimCLR) is bench.py, C4 (MAE) is tools/perf_probe.py vit."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200.core import ParamStore  # noqa: E402
from passl_b200.optimizer import AdamW, Momentum  # noqa: E402


def timeit(step, iters=5, warmup=2):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        step()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def moco_v2(B=256):
    from passl_b200.modeling import build_model
    from passl_b200.utils.config import get_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"))
    m = build_model(dict(cfg.model)).cuda()
    sq, _ = m.build_param_stores()
    opt = Momentum(sq, lr=0.03, momentum=0.9, weight_decay=1e-4)
    a, b = torch.randn(B, 3, 224, 224, device="cuda"), torch.randn(B, 3, 224, 224, device="cuda")

    def step():
        opt.clear_grad()
        out = m(a, b)
        out["loss"].backward()
        opt.step()
    ms = timeit(step)
    print("C3 MoCo v2 R50 K=65536 bs %d: %.2f ms/step -> %.0f img/s ; mem %.1f GB" % (B, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 2**30), flush=True)


def clip(B):
    from passl_b200.modeling import build_model
    arch = dict(name="CLIP", embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                pre_norm=True, proj=True, patch_bias=False, context_length=77, vocab_size=49408, transformer_width=512,
                transformer_heads=8, transformer_layers=12, qkv_bias=True)
    m = build_model(dict(name="CLIPWrapper", architecture=arch, head=dict(name="CLIPHead"))).cuda()
    with torch.no_grad():                       # the reference's (2*depth)x projection init overflows bf16 activations at depth 12
        for blk in m.model.text.blocks:
            blk.proj.weight.mul_(1.0 / 24)
            blk.fc2.weight.mul_(1.0 / 24)
    st = ParamStore(m)
    opt = AdamW(st, lr=1e-4, beta2=0.98, weight_decay=0.0005)
    img = torch.randn(B, 3, 224, 224, device="cuda")
    text = torch.randint(1, 49407, (B, 77), device="cuda")
    text[torch.arange(B), torch.randint(1, 77, (B,))] = 49407

    def step():
        opt.clear_grad()
        out = m(img, text)
        out["loss"].backward()
        opt.step()
    ms = timeit(step, iters=4)
    print("C5 CLIP ViT-B/16 + text 12x512 bs %d: %.2f ms/step -> %.0f pairs/s ; mem %.1f GB" % (B, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 2**30), flush=True)
    del m, st, opt
    torch.cuda.empty_cache()


def mocov3(B=256):
    from passl_b200.models import mocov3_vit_base_pretrain
    m = mocov3_vit_base_pretrain().cuda()
    st, _ = m.build_param_stores()
    opt = AdamW(st, lr=1.5e-4, weight_decay=0.1)
    a, b = torch.randn(B, 3, 224, 224, device="cuda"), torch.randn(B, 3, 224, 224, device="cuda")

    def step():
        opt.clear_grad()
        loss = m([a, b])
        loss.backward()
        opt.step()
    ms = timeit(step, iters=4)
    print("MoCo v3 ViT-B/16 bs %d (2 views, momentum encoder): %.2f ms/step -> %.0f img/s ; mem %.1f GB" % (B, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 2**30), flush=True)


if __name__ == "__main__":
    for fn, arg in ((moco_v2, 256), (clip, 256), (clip, 1024), (mocov3, 256)):
        t0 = time.time()
        try:
            fn(arg)
        except Exception as ex:  # report and continue with the next configuration
            print("%s(%s) failed: %r" % (fn.__name__, arg, ex), flush=True)
        torch.cuda.empty_cache()
