"""CPU restatement of the image input stage that feeds the path — TEST INFRASTRUCTURE, not a product path: only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

What the reference does per view (configs/simclr/simclr_r50_IM.yaml:35-61, passl_v110/datasets/imagenet.py:46-63):
PIL image -> RandomResizedCrop (crop, then PIL resize to 224 with bilinear / bicubic) -> [ColorJitter, GaussianBlur: not restated
here] -> RandomGrayscale -> RandomHorizontalFlip -> Transpose (HWC -> CHW ndarray) -> NormalizeImage (x * scale - mean) / std.

The pixel arithmetic of crop / resize / grayscale lives in Pillow (third-party; 12.2.0 in this image), whose 8-bit resampler
(src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) and
RGB -> L conversion (Convert.c: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16) are restated below in numpy.  Pinned:
tests/test_oracle_input_stage_cpu.py compares every function here with Pillow itself, bit for bit, on random images.
The crop-parameter draw follows the reference's in-repo sampler (transforms.py:517-557 get_params).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = np.abs(x)
    return np.where(x < 1.0, 1.0 - x, 0.0)


def _bicubic(x):
    a = -0.5
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1,
                    np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def precompute_coeffs(in_size, out_size, interpolation):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the box (0, in_size): -> (bounds int32 [out, 2] = (xmin, count),
    kk int32 [out, ksize] fixed-point weights)."""
    filt, fsupport = FILTERS[interpolation]
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)               # C cast: truncation toward zero
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        w = filt((x + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                                              # sequential sum, like the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, (-0.5 + w * (1 << PRECISION_BITS)), (0.5 + w * (1 << PRECISION_BITS)))
        kk[xx, :xmax] = fixed.astype(np.int64).astype(np.int32)  # (int) cast truncates toward zero
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_u8(img, out_h, out_w, interpolation="bilinear"):
    """PIL Image.resize((out_w, out_h), resample) of an HWC uint8 image: horizontal pass, uint8 intermediate, vertical pass; a pass
    is skipped when its size does not change (Resample.c ImagingResampleInner need_horizontal / need_vertical)."""
    h, w, _ = img.shape
    cur = img
    if out_w != w:
        bounds, kk = precompute_coeffs(w, out_w, interpolation)
        tmp = np.empty((h, out_w, img.shape[2]), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = (cur[:, x0:x0 + n, :].astype(np.int64) * kk[xx, :n].astype(np.int64)[None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
    if out_h != h:
        bounds, kk = precompute_coeffs(h, out_h, interpolation)
        out = np.empty((out_h, cur.shape[1], img.shape[2]), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = (cur[y0:y0 + n].astype(np.int64) * kk[yy, :n].astype(np.int64)[:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
            out[yy] = _clip8(acc)
        cur = out
    return cur.copy() if cur is img else cur


def resized_crop_u8(img, i, j, h, w, size, interpolation="bilinear"):
    """crop(j, i, j + w, i + h) then resize to (size, size): the RandomResizedCrop image op."""
    return resize_u8(np.ascontiguousarray(img[i:i + h, j:j + w]), size, size, interpolation)


def grayscale3_u8(img):
    """img.convert('L') replicated to three channels (RandomGrayscale with num_output_channels = 3, transforms.py:150-170)."""
    r, g, b = (img[..., c].astype(np.int64) for c in range(3))
    L = ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)
    return np.stack([L, L, L], axis=-1)


def hflip_u8(img):
    return np.ascontiguousarray(img[:, ::-1])


def transpose_normalize(img, scale=1.0 / 255.0, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """Transpose (HWC -> CHW) + NormalizeImage (transforms.py:462-467): uint8 * python float -> float64, then
    (x - float32(mean)) / float32(std) as paddle.vision's numpy normalize does, result stored as float32."""
    x = img.transpose(2, 0, 1).astype(np.float64) * scale
    m = np.float32(np.array(mean).reshape(-1, 1, 1)).astype(np.float64)
    s = np.float32(np.array(std).reshape(-1, 1, 1)).astype(np.float64)
    return ((x - m) / s).astype(np.float32)


def get_params(width, height, scale, ratio, rng):
    """Crop-box draw of transforms.py:517-557, -> (top, left, crop_h, crop_w).  Up to ten draws of an area fraction in `scale` and
    an aspect ratio log-uniform in `ratio`; the first box that fits wins and is placed uniformly.  Otherwise the largest centred box
    whose aspect ratio lies inside `ratio`.  `rng`: uniform(a, b) / randint(a, b) with Python `random` semantics (inclusive ends)."""
    lo, hi = math.log(ratio[0]), math.log(ratio[1])
    for _ in range(10):
        box_area = rng.uniform(scale[0], scale[1]) * (width * height)
        aspect = math.exp(rng.uniform(lo, hi))
        cw, ch = int(round(math.sqrt(box_area * aspect))), int(round(math.sqrt(box_area / aspect)))
        if cw <= width and ch <= height:
            top = rng.randint(0, height - ch)
            return top, rng.randint(0, width - cw), ch, cw
    image_aspect = width / height
    if image_aspect < min(ratio):
        cw, ch = width, int(round(width / min(ratio)))
    elif image_aspect > max(ratio):
        cw, ch = int(round(height * max(ratio))), height
    else:
        cw, ch = width, height
    return (height - ch) // 2, (width - cw) // 2, ch, cw


# ---------------------------------------------------------------------------------------------------------------------------------
# ColorJitter pixel arithmetic (configs/simclr/simclr_r50_IM.yaml:41-48: brightness 0.4, contrast 0.4, saturation 0.4, hue 0.1).
# paddle.vision's PIL backend calls ImageEnhance.{Brightness, Contrast, Color}(img).enhance(factor) and, for hue, shifts the H plane
# of img.convert('HSV') by uint8(hue_factor * 255) with wrap-around and converts back.  Restated from Pillow (libImaging/Blend.c,
# Convert.c rgb2hsv / hsv2rgb, ImageEnhance.py, ImageStat.py) and pinned bit-exact against Pillow: the HSV conversions over all 2^24
# colours, the blend over all (a, b) byte pairs for hundreds of factors (tests/test_oracle_input_stage_cpu.py).
# ---------------------------------------------------------------------------------------------------------------------------------
def blend_u8(degenerate, img, alpha):
    """Image.blend(degenerate, img, alpha): float32 a + alpha * (b - a), truncated; clipped to [0, 255] when alpha is outside [0, 1]."""
    al = np.float32(alpha)
    a = degenerate.astype(np.float32)
    t = a + al * (img.astype(np.int32) - degenerate.astype(np.int32)).astype(np.float32)
    if 0 <= alpha <= 1.0:
        return t.astype(np.int64).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255.0, 255, t.astype(np.int64))).astype(np.uint8)


def luma_u8(img):
    r, g, b = (img[..., c].astype(np.int64) for c in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def adjust_brightness(img, factor):
    return blend_u8(np.zeros_like(img), img, factor)


def adjust_contrast(img, factor):
    L = luma_u8(img)
    mean = int(float(L.astype(np.int64).sum()) / L.size + 0.5)                  # ImageStat mean of the L image, rounded half up
    return blend_u8(np.full_like(img, mean), img, factor)


def adjust_saturation(img, factor):
    L = luma_u8(img)
    return blend_u8(np.stack([L, L, L], axis=-1), img, factor)


def rgb_to_hsv_u8(img):
    """Convert.c rgb2hsv: channel ratios in float32, hue assembled in double and rounded to float32, scaled by 255.0 in double."""
    f32, f64 = np.float32, np.float64
    r, g, b = (img[..., c].astype(np.int32) for c in range(3))
    maxc, minc = np.maximum(r, np.maximum(g, b)), np.minimum(r, np.minimum(g, b))
    with np.errstate(divide="ignore", invalid="ignore"):
        cr = (maxc - minc).astype(f32)
        s = cr / maxc.astype(f32)
        rc, gc, bc = ((maxc - c).astype(f32) / cr for c in (r, g, b))
        h = np.where(r == maxc, bc.astype(f64) - gc.astype(f64),
                     np.where(g == maxc, 2.0 + rc.astype(f64) - bc.astype(f64), 4.0 + gc.astype(f64) - rc.astype(f64))).astype(f32)
        h = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32)
        uh = np.clip((h.astype(f64) * 255.0).astype(np.int64), 0, 255)
        us = np.clip((s.astype(f64) * 255.0).astype(np.int64), 0, 255)
    gray = maxc == minc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], axis=-1).astype(np.uint8)


def hsv_to_rgb_u8(hsv):
    """Convert.c hsv2rgb (colorsys sextants): p, q, t = round(v (1 - s..)), C round() = half away from zero."""
    f32, f64 = np.float32, np.float64
    h, s, v = (hsv[..., c].astype(np.int32) for c in range(3))
    hd = h.astype(f64) * 6.0 / 255.0
    i = np.floor(hd).astype(np.int64)
    f = (hd - i).astype(f32).astype(f64)
    fs = (s.astype(f64) / 255.0).astype(f32).astype(f64)
    p, q, t = (np.clip(np.floor(v * x + 0.5), 0, 255).astype(np.int64) for x in (1.0 - fs, 1.0 - fs * f, 1.0 - fs * (1.0 - f)))
    k = i % 6
    out = np.stack([np.choose(k, [v, q, p, p, t, v]), np.choose(k, [t, v, v, q, p, p]), np.choose(k, [p, p, t, v, v, q])], axis=-1)
    return np.where((s == 0)[..., None], np.stack([v, v, v], axis=-1), out).astype(np.uint8)


def adjust_hue(img, hue_factor):
    """H plane += uint8(hue_factor * 255) modulo 256 (numpy uint8 cast of the product: truncation toward zero, then wrap)."""
    hsv = rgb_to_hsv_u8(img)
    shift = np.array(hue_factor * 255).astype(np.int64).astype(np.uint8)       # np.uint8(float): C cast, negative values wrap
    hsv[..., 0] = hsv[..., 0] + shift
    return hsv_to_rgb_u8(hsv)


# ---------------------------------------------------------------------------------------------------------------------------------
# GaussianBlur (transforms.py:173-191, default path): cv2.GaussianBlur(np.array(x), (23, 23), sigma), sigma ~ np.random.uniform.
# For uint8 images OpenCV runs its "bit-exact" fixed-point filter (imgproc/src/smooth.dispatch.cpp, smooth.simd.hpp,
# fixedpoint.inl.hpp; OpenCV 4.13 in this image): the normalised Gaussian taps are turned into 8.8 fixed point by error diffusion
# from the edge inwards with the centre tap taking the remainder (taps sum to exactly 256), the horizontal pass keeps 8.8 results
# unrounded, the vertical pass rounds once: (sum + 2^15) >> 16; borders are BORDER_REFLECT_101.  Restated below and pinned against
# cv2.GaussianBlur itself over thousands of sigmas in [0.1, 2] (tests/test_oracle_input_stage_cpu.py).
# ---------------------------------------------------------------------------------------------------------------------------------
def gaussian_taps_fixed(ksize, sigma, bits=8):
    half = (ksize - 1) // 2
    scale = -0.5 / (sigma * sigma)
    vals = [math.exp(scale * float((i - half) ** 2)) for i in range(half)]
    norm = 1.0 / (2.0 * sum(vals) + 1.0)
    taps, err, total = [0] * ksize, 0.0, 0
    for i in range(half):
        adj = vals[i] * norm * float(1 << bits) + err
        q = int(round(adj))                                      # cvRound: ties to even, like Python's round()
        err = adj - q
        taps[i] = taps[ksize - 1 - i] = q
        total += q
    taps[half] = (1 << bits) - 2 * total
    return np.array(taps, dtype=np.int64)


def gaussian_blur_u8(img, ksize, sigma):
    k = gaussian_taps_fixed(ksize, sigma)
    r = ksize // 2
    H, W = img.shape[:2]
    pad = np.pad(img.astype(np.int64), ((r, r), (r, r), (0, 0)), mode="reflect")      # numpy 'reflect' = BORDER_REFLECT_101
    hres = sum(k[i] * pad[:, i:i + W] for i in range(ksize))
    out = sum(k[j] * hres[j:j + H] for j in range(ksize))
    return ((out + (1 << 15)) >> 16).astype(np.uint8)
