#!/usr/bin/env python
"""Times the device input stage (csrc/input_stage.cu) kernel group by kernel group with CUDA events on the launching stream and
prints achieved GB/s against the algorithmic bytes of DESIGN.md §3 — first measurement of SURVEY.md §8 f-2 (not a bench line).

    python tools/input_stage_probe.py [--batch 512] [--size 224] [--iters 20]
"""
import argparse
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from passl_b200.data import (ImageBatch, SyntheticDecodedImages, TwoViewInputStage, color_jitter_u8, gaussian_blur_u8,  # noqa: E402
                             resized_crop_u8, views_finalize)


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    images = SyntheticDecodedImages(args.batch, 1, dev).images
    batch = ImageBatch(images)
    stage = TwoViewInputStage(size=args.size, scale=(0.1, 1.0), interpolation="bicubic", jitter_p=0.8, gray_p=0.2, blur_p=0.5,
                              rng=random.Random(0), np_rng=np.random.RandomState(0))
    item_img, item_box, gray, flip, plans = stage.draw(batch)
    sigmas = stage.last_sigmas
    M, S = len(item_img), args.size
    crop_bytes = sum(3 * h * w for (_, _, h, w) in item_box)
    tmp_bytes = sum(3 * h * S for (_, _, h, w) in item_box)
    out_u8 = 3 * S * S * M
    u8 = resized_crop_u8(batch, item_img, item_box, S, "bicubic")
    rows = []
    t = timed(lambda: resized_crop_u8(batch, item_img, item_box, S, "bicubic", check=False), args.iters)
    rows.append(("crop + resize (coeffs, h, v)", t, crop_bytes + 2 * tmp_bytes + out_u8))
    n_j = sum(1 for p in plans if p)
    t = timed(lambda: color_jitter_u8(u8.clone(), plans), args.iters) - timed(lambda: u8.clone(), args.iters)
    rows.append(("colour jitter (%d of %d views, 4 ops)" % (n_j, M), t, n_j * 3 * S * S * (4 * 2 + 1)))
    n_b = sum(1 for sg in sigmas if sg is not None)
    t = timed(lambda: gaussian_blur_u8(u8.clone(), sigmas), args.iters) - timed(lambda: u8.clone(), args.iters)
    rows.append(("gaussian blur 23x23 (%d views)" % n_b, t, n_b * 3 * S * S * (1 + 2 + 2 + 1)))
    t = timed(lambda: views_finalize(u8, gray, flip), args.iters)
    rows.append(("finalize (gray, flip, CHW, normalise)", t, out_u8 + 4 * out_u8))
    t_all = timed(lambda: stage(batch), max(3, args.iters // 4))
    print("input stage, %d images -> %d views of %dx%d (host draw + small H2D included in 'whole stage')" % (args.batch, M, S, S))
    for name, ms, nbytes in rows:
        print("  %-42s %8.3f ms   %7.1f GB/s   (%.1f MB algorithmic)" % (name, ms, nbytes / ms / 1e6, nbytes / 1e6))
    print("  %-42s %8.3f ms   %7.0f images/s" % ("whole stage (two views per image)", t_all, args.batch / t_all * 1e3))


if __name__ == "__main__":
    main()
