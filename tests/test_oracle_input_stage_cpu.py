"""oracle/input_stage.py pinned: the resize / grayscale arithmetic bit for bit against Pillow itself (the reference's image backend,
12.x in this image), the crop-box draw against boxes produced by the reference's own sampler
(tests/golden/make_golden_input.py -> reference_crop_params.npz), the float stage against its definition."""
import os
import random

import numpy as np
import pytest

import oracle.input_stage as O

Image = pytest.importorskip("PIL.Image")
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_crop_params.npz"))


@pytest.mark.parametrize("interp", ["bilinear", "bicubic"])
@pytest.mark.parametrize("shape", [(300, 400, 224, 224), (100, 80, 224, 224), (500, 333, 224, 224), (224, 500, 224, 224),
                                   (37, 41, 64, 64), (640, 480, 32, 32), (225, 223, 224, 224), (1, 1, 8, 8), (224, 224, 224, 224)])
def test_resize_is_pillow_bit_exact(interp, shape):
    H, W, oh, ow = shape
    rng = np.random.RandomState(H * 1000 + W)
    img = rng.randint(0, 256, size=(H, W, 3)).astype(np.uint8)
    if H > 100:
        img[: H // 2] = (img[: H // 2] // 128) * 255                 # hard edges: exercises the negative bicubic lobes and the clip
    want = np.asarray(Image.fromarray(img).resize((ow, oh), getattr(Image, interp.upper())))
    assert np.array_equal(O.resize_u8(img, oh, ow, interp), want)


def test_resized_crop_and_grayscale_and_flip_match_pillow():
    rng = np.random.RandomState(7)
    img = rng.randint(0, 256, size=(375, 500, 3)).astype(np.uint8)
    pil = Image.fromarray(img)
    for (i, j, h, w) in [(10, 20, 200, 300), (0, 0, 375, 500), (300, 400, 75, 100), (100, 100, 1, 1)]:
        for interp in ("bilinear", "bicubic"):
            want = np.asarray(pil.crop((j, i, j + w, i + h)).resize((224, 224), getattr(Image, interp.upper())))
            assert np.array_equal(O.resized_crop_u8(img, i, j, h, w, 224, interp), want)
    gray = np.asarray(pil.convert("L"))
    assert np.array_equal(O.grayscale3_u8(img), np.stack([gray] * 3, axis=-1))
    assert np.array_equal(O.hflip_u8(img), np.asarray(pil.transpose(Image.FLIP_LEFT_RIGHT)))


def test_crop_boxes_equal_the_reference_sampler():
    for tag in "abcdef":
        W, H, s0, s1, r0, r1, seed = G["args_" + tag]
        random.seed(int(seed))
        got = [O.get_params(int(W), int(H), (s0, s1), (r0, r1), random) for _ in range(64)]
        assert np.array_equal(np.array(got), G["boxes_" + tag]), tag
    assert len({tuple(b) for b in G["boxes_d"]}) == 1 and len({tuple(b) for b in G["boxes_a"]}) > 32   # d: always the fallback box


def test_transpose_normalize_definition():
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, size=(5, 7, 3)).astype(np.uint8)
    out = O.transpose_normalize(img)
    assert out.dtype == np.float32 and out.shape == (3, 5, 7)
    mean, std = np.float32([0.485, 0.456, 0.406]), np.float32([0.229, 0.224, 0.225])
    for c in range(3):
        want = ((img[..., c].astype(np.float64) * (1.0 / 255.0) - np.float64(mean[c])) / np.float64(std[c])).astype(np.float32)
        assert np.array_equal(out[c], want)


# ------------------------------------------------------------------ ColorJitter arithmetic vs Pillow -------------------------------
def test_hsv_conversions_are_pillow_exact_over_all_colours():
    v = np.arange(1 << 24, dtype=np.uint32)
    cube = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(O.rgb_to_hsv_u8(cube), np.asarray(Image.fromarray(cube, "RGB").convert("HSV")))
    assert np.array_equal(O.hsv_to_rgb_u8(cube), np.asarray(Image.fromarray(cube, "HSV").convert("RGB")))


def test_blend_is_pillow_exact_over_all_byte_pairs():
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    b = np.ascontiguousarray(a.T)
    rng = random.Random(0)
    for alpha in [0.0, 1.0, 0.5, 0.6, 1.4, 1.7999, 0.2000001, 1e-3, 2.0, 1.0000001] + [rng.uniform(0.0, 2.0) for _ in range(200)]:
        want = np.asarray(Image.blend(Image.fromarray(a, "L"), Image.fromarray(b, "L"), alpha))
        assert np.array_equal(O.blend_u8(a, b, alpha), want), alpha


def test_colour_jitter_ops_match_pillow_enhancers():
    from PIL import ImageEnhance
    rng = np.random.RandomState(11)
    r = random.Random(5)
    for shape in [(64, 64), (37, 91)]:
        img = rng.randint(0, 256, size=shape + (3,)).astype(np.uint8)
        img[: shape[0] // 3] //= 4                                                 # a dark band: moves the contrast mean
        pil = Image.fromarray(img)
        for _ in range(12):
            f = r.uniform(0.6, 1.4)
            assert np.array_equal(O.adjust_brightness(img, f), np.asarray(ImageEnhance.Brightness(pil).enhance(f)))
            assert np.array_equal(O.adjust_contrast(img, f), np.asarray(ImageEnhance.Contrast(pil).enhance(f)))
            assert np.array_equal(O.adjust_saturation(img, f), np.asarray(ImageEnhance.Color(pil).enhance(f)))
            hf = r.uniform(-0.1, 0.1)
            h, s, v = pil.convert("HSV").split()                                   # the PIL path of paddle's adjust_hue
            np_h = np.array(h, dtype=np.uint8)
            with np.errstate(over="ignore"):
                np_h += np.array(hf * 255).astype(np.uint8)
            want = Image.merge("HSV", (Image.fromarray(np_h, "L"), s, v)).convert("RGB")
            assert np.array_equal(O.adjust_hue(img, hf), np.asarray(want)), hf


def test_gaussian_blur_is_opencv_bit_exact():
    cv2 = pytest.importorskip("cv2")
    rng, r = np.random.RandomState(0), random.Random(0)
    sigmas = [0.1, 2.0, 0.1000001, 1.9999999, 0.5, 1.0, 1.5, 0.25, 0.75, 0.3, 1.2] + [r.uniform(0.1, 2.0) for _ in range(1500)]
    for sigma in sigmas:
        img = rng.randint(0, 256, size=(24, 25, 3)).astype(np.uint8)
        assert np.array_equal(O.gaussian_blur_u8(img, 23, sigma), cv2.GaussianBlur(img, (23, 23), sigma)), repr(sigma)
    img = rng.randint(0, 256, size=(224, 224, 3)).astype(np.uint8)
    img[:100] = (img[:100] // 128) * 255
    for sigma in (0.1, 0.77, 2.0):
        assert np.array_equal(O.gaussian_blur_u8(img, 23, sigma), cv2.GaussianBlur(img, (23, 23), sigma))
    for sigma in (0.1, 0.9, 2.0):
        taps = O.gaussian_taps_fixed(23, sigma)
        assert taps.sum() == 256 and (taps >= 0).all() and np.array_equal(taps, taps[::-1])
