"""The image input stage through the C ABI on the GPU (csrc/input_stage.cu): the cases of tests/test_input_stage_host_cpu.py — ragged
source images, edge boxes, both filters — bit-exact against Pillow / the oracle, the float stage bit-exact against the oracle, and
the two-view stage end to end.  (Named to sort last: it is the newest row of SURVEY.md §8, f-2.)"""
import os
import random
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("S", [224, 32])
def test_resized_crop_and_finalize_match_the_oracle(interp, S):
    import oracle.input_stage as O
    from test_input_stage_host_cpu import make_cases
    from passl_b200.data import ImageBatch, resized_crop_u8, views_finalize
    c = make_cases()
    batch = ImageBatch(c["images"])
    boxes = [tuple(int(v) for v in b) for b in c["item_box"]]
    idx = [int(i) for i in c["item_img"]]
    u8 = resized_crop_u8(batch, idx, boxes, S, interp)
    got = u8.cpu().numpy()
    for m, (n, (i, j, h, w)) in enumerate(zip(idx, boxes)):
        assert np.array_equal(got[m], O.resized_crop_u8(c["images"][n], i, j, h, w, S, interp)), (m, interp)
    gray = [0, 1, 0, 1, 0, 0, 1, 0, 1]
    flip = [0, 0, 1, 1, 0, 1, 0, 1, 0]
    out = views_finalize(u8, gray, flip).cpu().numpy()
    for m in range(len(idx)):
        img = got[m]
        if gray[m]:
            img = O.grayscale3_u8(img)
        if flip[m]:
            img = O.hflip_u8(img)
        assert np.array_equal(out[m], O.transpose_normalize(img)), m


def test_bad_box_is_refused_on_the_host_and_flagged_on_the_device():
    from test_input_stage_host_cpu import make_cases
    from passl_b200 import _lib
    from passl_b200.data import ImageBatch, resized_crop_u8
    c = make_cases()
    batch = ImageBatch(c["images"])
    with pytest.raises(ValueError):
        resized_crop_u8(batch, [0], [(300, 400, 76, 100)], 32)
    # the device-side guard, reached by lying to the host check about the image size
    batch.heights[0] += 1
    with pytest.raises(_lib.PasslB200Error):
        resized_crop_u8(batch, [0, 1], [(300, 400, 76, 100), (0, 0, 64, 48)], 32)


def test_two_view_stage_end_to_end():
    """N images -> (view_1, view_2); every view equals the oracle applied with the decisions the stage drew."""
    import oracle.input_stage as O
    from test_input_stage_host_cpu import oracle_jitter
    from passl_b200.data import ImageBatch, TwoViewInputStage
    rng = np.random.RandomState(5)
    images = [rng.randint(0, 256, size=(int(h), int(w), 3)).astype(np.uint8) for h, w in [(240, 320), (333, 250), (128, 128), (96, 400)]]
    kw = dict(size=64, scale=(0.2, 1.0), interpolation="bicubic", jitter_p=0.6, gray_p=0.5, blur_p=0.5)
    stage = TwoViewInputStage(rng=random.Random(11), np_rng=np.random.RandomState(4), **kw)
    twin = TwoViewInputStage(rng=random.Random(11), np_rng=np.random.RandomState(4), **kw)
    batch = ImageBatch(images)
    v1, v2 = stage(batch)
    item_img, item_box, gray, flip, plans = twin.draw(batch)
    assert any(plans) and not all(plans)                                            # some views jittered, some not (jitter_p 0.6)
    sig = twin.last_sigmas
    assert any(s_ is not None and g for s_, g in zip(sig, gray)) and any(s_ is None and g for s_, g in zip(sig, gray))
    assert v1.shape == (4, 3, 64, 64) and v2.shape == (4, 3, 64, 64) and v1.dtype == torch.float32
    both = torch.cat([v1, v2]).cpu().numpy()
    for m, (n, (i, j, h, w)) in enumerate(zip(item_img, item_box)):
        img = oracle_jitter(O.resized_crop_u8(images[n], i, j, h, w, 64, "bicubic"), plans[m])
        if gray[m]:
            img = O.grayscale3_u8(img)
        if twin.last_sigmas[m] is not None:                                       # pipeline order: jitter, grayscale, blur, flip
            img = O.gaussian_blur_u8(img, 23, twin.last_sigmas[m])
        if flip[m]:
            img = O.hflip_u8(img)
        assert np.array_equal(both[m], O.transpose_normalize(img)), m
    assert not np.array_equal(both[0], both[4])                                   # two different views of sample 0


def test_colour_jitter_kernels_match_the_oracle():
    """ColorJitter through the C ABI: shuffled op orders, contrast twice (mean re-taken), extrapolating blend, flat grey view."""
    from test_input_stage_host_cpu import jitter_cases, oracle_jitter
    from passl_b200.data import color_jitter_u8
    img, ops, factors, plan = jitter_cases()
    dev = torch.from_numpy(img.copy()).cuda()
    got = color_jitter_u8(dev, [[(op, f) for op, f in row if op] for row in plan]).cpu().numpy()
    for m in range(img.shape[0]):
        assert np.array_equal(got[m], oracle_jitter(img[m], [(op, f) for op, f in plan[m] if op])), (m, plan[m])
    assert not np.array_equal(got[0], img[0])
    big = torch.randint(0, 256, (3, 224, 224, 3), dtype=torch.uint8, device="cuda")   # several blocks per view: exact luma sums
    ref = big.cpu().numpy()
    out = color_jitter_u8(big.clone(), [[(2, 1.3)], [], [(3, 0.7), (2, 0.8)]]).cpu().numpy()
    assert np.array_equal(out[1], ref[1])
    assert np.array_equal(out[0], oracle_jitter(ref[0], [(2, 1.3)])) and np.array_equal(out[2], oracle_jitter(ref[2], [(3, 0.7), (2, 0.8)]))


def test_gaussian_blur_kernels_match_the_oracle():
    """cv2.GaussianBlur's fixed-point uint8 path through the C ABI: several sigmas, an untouched view, reflect-101 borders."""
    import oracle.input_stage as O
    from passl_b200.data import gaussian_blur_u8
    rng = np.random.RandomState(6)
    for S in (24, 224):
        img = rng.randint(0, 256, size=(4, S, S, 3)).astype(np.uint8)
        img[1, : S // 2] = (img[1, : S // 2] // 128) * 255
        sigmas = [0.1, 2.0, None, 0.8371]
        got = gaussian_blur_u8(torch.from_numpy(img.copy()).cuda(), sigmas).cpu().numpy()
        for m, sg in enumerate(sigmas):
            want = img[m] if sg is None else O.gaussian_blur_u8(img[m], 23, sg)
            assert np.array_equal(got[m], want), (S, m)


def test_trainer_with_the_device_input_stage():
    """configs/moco + `dataloader.train.device_input_stage=True`: uint8 images -> device stage (recipe from the YAML) -> MoCo steps."""
    from passl_b200.engine.trainer import Trainer
    from passl_b200.utils.config import get_config
    root = os.path.dirname(HERE)
    cfg = get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"),
                     ["model.K=1024", "dataloader.train.sampler.batch_size=16", "dataloader.train.device_input_stage=True",
                      "dataloader.train.dataset.transforms.0.size=64", "total_iters=3", "epochs=1", "log_config.interval=100"])
    tr = Trainer(cfg)
    assert type(tr.dataloader).__name__ == "DeviceAugmentedTwoViews" and tr.dataloader.stage.size == 64
    v1, v2 = next(iter(tr.dataloader))
    assert v1.shape == (16, 3, 64, 64) and v1.dtype == torch.float32 and torch.isfinite(v1).all() and not torch.equal(v1, v2)
    assert abs(v1.mean().item()) < 1.5 and 0.01 < v1.std().item() < 3.0            # normalised pixel statistics of resampled noise
    out = tr.train()
    assert np.isfinite(float(out["loss"].detach()))
