"""Image input stage on the device: decoded uint8 HWC images in HBM -> two augmented, normalised NCHW views.

Reference (CPU workers, per sample; passl_v110/datasets/imagenet.py:46-63 with configs/simclr/simclr_r50_IM.yaml:35-83):
`sample1 = transform(sample); sample2 = transform(sample)` with transform = RandomResizedCrop, then per view
`RandomApply(ColorJitter) -> RandomGrayscale -> RandomApply(GaussianBlur) -> RandomHorizontalFlip -> Transpose -> NormalizeImage`.

Built here: the crop-box draw on the host (same algorithm and `random` call order as transforms.py:517-557), then on the GPU
(csrc/input_stage.cu) crop + Pillow-exact resize, ColorJitter (Pillow's ImageEnhance blends and HSV hue shift), grayscale, flip,
CHW + normalise — every pixel op bit-exact against Pillow.  GaussianBlur (cv2, 23x23) is not built, and ColorJitter's use of
the `random` stream restates PaddlePaddle code that is not in the reference tree; so pixel arithmetic is pinned, the exact
sequence of random draws of a full reference worker is not.
"""
import math
import random

import torch

from .. import _lib

INTERPOLATION = {"bilinear": 0, "bicubic": 1}


def random_resized_crop_params(width, height, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.), rng=random):
    """(top, left, crop_h, crop_w) like RandomResizedCrop.get_params (transforms.py:517-557): up to ten draws of an area fraction
    and a log-uniform aspect ratio, first box that fits placed uniformly; else the largest centred box within the ratio range."""
    log_lo, log_hi = math.log(ratio[0]), math.log(ratio[1])
    for _ in range(10):
        target = rng.uniform(scale[0], scale[1]) * (width * height)
        aspect = math.exp(rng.uniform(log_lo, log_hi))
        cw = int(round(math.sqrt(target * aspect)))
        ch = int(round(math.sqrt(target / aspect)))
        if cw <= width and ch <= height:
            top = rng.randint(0, height - ch)
            left = rng.randint(0, width - cw)
            return top, left, ch, cw
    whole = width / height
    if whole < min(ratio):
        cw, ch = width, int(round(width / min(ratio)))
    elif whole > max(ratio):
        cw, ch = int(round(height * max(ratio))), height
    else:
        cw, ch = width, height
    return (height - ch) // 2, (width - cw) // 2, ch, cw


class ImageBatch:
    """Decoded RGB images of different sizes packed back to back in one uint8 device buffer."""

    def __init__(self, images, device=None):
        """images: sequence of uint8 [H, W, 3] tensors / arrays (host or device)."""
        ts = [torch.as_tensor(im) for im in images]
        for t in ts:
            if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
                raise ValueError("images must be uint8 [H, W, 3], got %s %s" % (t.dtype, tuple(t.shape)))
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.heights = [int(t.shape[0]) for t in ts]
        self.widths = [int(t.shape[1]) for t in ts]
        sizes = [h * w * 3 for h, w in zip(self.heights, self.widths)]
        offs = [0]
        for s in sizes[:-1]:
            offs.append(offs[-1] + s)
        self.data = torch.cat([t.reshape(-1).to(device, non_blocking=True) for t in ts])
        self.src_off = torch.tensor(offs, dtype=torch.int64, device=device)
        self.src_h = torch.tensor(self.heights, dtype=torch.int32, device=device)
        self.src_w = torch.tensor(self.widths, dtype=torch.int32, device=device)
        self.device = device

    def __len__(self):
        return len(self.heights)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t):
    if not t.is_cuda:
        raise _lib.PasslB200Error("passl_b200 kernels need CUDA tensors (there is no CPU fallback)")


def resized_crop_u8(batch, item_img, item_box, size=224, interpolation="bilinear", check=True):
    """Crop + PIL-exact resize of `len(item_img)` views.  item_img: source index per view, item_box: (top, left, h, w) per view
    (host lists).  -> uint8 [items, size, size, 3] on the device."""
    _need_cuda(batch.data)
    lib = _lib.load()
    items = len(item_img)
    for n, (t, l, h, w) in zip(item_img, item_box):
        if not (0 <= n < len(batch)) or h <= 0 or w <= 0 or t < 0 or l < 0 or t + h > batch.heights[n] or l + w > batch.widths[n]:
            raise ValueError("crop box (top=%d, left=%d, h=%d, w=%d) does not lie inside image %d" % (t, l, h, w, n))
    if interpolation not in INTERPOLATION:
        raise NotImplementedError("interpolation %r (built: %s)" % (interpolation, sorted(INTERPOLATION)))
    inter = INTERPOLATION[interpolation]
    max_h = max(b[2] for b in item_box)
    kmax = lib.passl_b200_resample_kmax(max(max(b[2], b[3]) for b in item_box), size, inter)
    dev = batch.device
    d_img = torch.tensor(list(item_img), dtype=torch.int32, device=dev)
    d_box = torch.tensor([list(b) for b in item_box], dtype=torch.int32, device=dev).reshape(-1)
    nbytes = lib.passl_b200_resized_crop_workspace_bytes(items, size, max_h, kmax)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty((items, size, size, 3), dtype=torch.uint8, device=dev)
    _lib.check(lib.passl_b200_resized_crop_u8(batch.data.data_ptr(), batch.src_off.data_ptr(), batch.src_h.data_ptr(),
                                              batch.src_w.data_ptr(), d_img.data_ptr(), d_box.data_ptr(), out.data_ptr(),
                                              ws.data_ptr(), nbytes, items, size, max_h, kmax, inter, _stream()), "resized_crop_u8")
    if check:                                            # one int back from the device: bad boxes / short tap table
        status = int(ws[:4].view(torch.int32).item())
        if status:
            raise _lib.PasslB200Error("resized_crop_u8: device status %d (1 = box outside its image, 2 = kmax too small)" % status)
    return out


def views_finalize(views_u8, gray, flip, scale=1.0 / 255.0, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """uint8 [items, S, S, 3] -> fp32 [items, 3, S, S]: grayscale where gray[m], mirror where flip[m], (x * scale - mean) / std."""
    import ctypes
    _need_cuda(views_u8)
    lib = _lib.load()
    items, S = views_u8.shape[0], views_u8.shape[1]
    assert views_u8.dtype == torch.uint8 and views_u8.shape == (items, S, S, 3) and views_u8.is_contiguous()
    dev = views_u8.device
    d_gray = torch.tensor([int(bool(g)) for g in gray], dtype=torch.int32, device=dev)
    d_flip = torch.tensor([int(bool(f)) for f in flip], dtype=torch.int32, device=dev)
    assert d_gray.numel() == items and d_flip.numel() == items
    out = torch.empty((items, 3, S, S), dtype=torch.float32, device=dev)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    _lib.check(lib.passl_b200_views_finalize_f32(views_u8.data_ptr(), d_gray.data_ptr(), d_flip.data_ptr(), out.data_ptr(), items, S,
                                                 float(scale), ctypes.cast(m3, ctypes.c_void_p), ctypes.cast(s3, ctypes.c_void_p),
                                                 _stream()), "views_finalize_f32")
    return out


JITTER_OPS = {"brightness": 1, "contrast": 2, "saturation": 3, "hue": 4}


def color_jitter_plan(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1, rng=random):
    """One ColorJitter application -> [(op, factor)] in execution order.

    Restates paddle.vision.transforms.ColorJitter (PaddlePaddle, not in the reference tree, so the consumption of the `random`
    stream below is unpinned): the four single-op transforms are put in a list, `random.shuffle`d, and each draws its factor with
    `random.uniform` when it runs — brightness / contrast / saturation from [max(0, 1 - v), 1 + v], hue from [-v, v]; an amount of
    0 keeps its slot in the shuffle but draws nothing."""
    slots = [("brightness", brightness), ("contrast", contrast), ("saturation", saturation), ("hue", hue)]
    rng.shuffle(slots)
    plan = []
    for name, amount in slots:
        if not amount:
            continue
        lo, hi = (-amount, amount) if name == "hue" else (max(0.0, 1.0 - amount), 1.0 + amount)
        plan.append((JITTER_OPS[name], rng.uniform(lo, hi)))
    return plan


def color_jitter_u8(views_u8, plans):
    """In-place ColorJitter of uint8 [items, S, S, 3] views; plans[m] = [(op, factor)] (at most four) or [] for an untouched view.
    Blend factors are passed as C floats like Pillow receives them; the hue shift is uint8(hue_factor * 255) computed here."""
    _need_cuda(views_u8)
    lib = _lib.load()
    items, S = views_u8.shape[0], views_u8.shape[1]
    assert views_u8.dtype == torch.uint8 and views_u8.shape == (items, S, S, 3) and views_u8.is_contiguous() and len(plans) == items
    ops, factors, contrast_positions = [], [], 0
    for plan in plans:
        if len(plan) > 4:
            raise ValueError("at most four jitter ops per view")
        row_o, row_f = [0, 0, 0, 0], [0.0, 0.0, 0.0, 0.0]
        for pos, (op, f) in enumerate(plan):
            row_o[pos] = int(op)
            row_f[pos] = float(int(f * 255) & 255) if op == JITTER_OPS["hue"] else float(f)
            if op == JITTER_OPS["contrast"]:
                contrast_positions |= 1 << pos
        ops.append(row_o)
        factors.append(row_f)
    if not any(any(r) for r in ops):
        return views_u8
    dev = views_u8.device
    d_ops = torch.tensor(ops, dtype=torch.int32, device=dev)
    d_fac = torch.tensor(factors, dtype=torch.float32, device=dev)
    ws = torch.empty(8 * items, dtype=torch.uint8, device=dev)
    _lib.check(lib.passl_b200_color_jitter_u8(views_u8.data_ptr(), d_ops.data_ptr(), d_fac.data_ptr(), ws.data_ptr(), ws.numel(), items, S,
                                              contrast_positions, _stream()), "color_jitter_u8")
    return views_u8


def grayscale_u8(views_u8, flags):
    """In-place RandomGrayscale (convert('L') replicated) of the flagged views — used when a grey view is blurred afterwards; views
    that are not blurred get their grayscale in views_finalize."""
    return color_jitter_u8(views_u8, [[(5, 0.0)] if f else [] for f in flags])


def gaussian_taps_fixed(ksize, sigma):
    """The 8.8 fixed-point tap row OpenCV's uint8 GaussianBlur uses for (ksize, sigma): normalised exp(-x^2 / 2 sigma^2) taps scaled
    by 256 and rounded by error diffusion from the edge inwards (ties to even), the centre tap taking the remainder so that the row
    sums to exactly 256 (imgproc/src/smooth.dispatch.cpp getGaussianKernelFixedPoint_ED).  Checked against cv2.GaussianBlur in the tests."""
    if ksize < 1 or ksize % 2 == 0 or not sigma > 0:
        raise ValueError("ksize must be odd and sigma positive, got %r, %r" % (ksize, sigma))
    half = (ksize - 1) // 2
    scale = -0.5 / (sigma * sigma)
    weights = [math.exp(scale * float((i - half) ** 2)) for i in range(half)]
    norm = 1.0 / (2.0 * sum(weights) + 1.0)
    taps, carry, used = [0] * ksize, 0.0, 0
    for i, w in enumerate(weights):
        exact = w * norm * 256.0 + carry
        q = int(round(exact))
        carry = exact - q
        taps[i] = taps[ksize - 1 - i] = q
        used += q
    taps[half] = 256 - 2 * used
    return taps


def gaussian_blur_u8(views_u8, sigmas, ksize=23):
    """In-place cv2-exact GaussianBlur of uint8 [items, S, S, 3] views; sigmas[m] = None leaves view m untouched."""
    _need_cuda(views_u8)
    lib = _lib.load()
    items, S = views_u8.shape[0], views_u8.shape[1]
    assert views_u8.dtype == torch.uint8 and views_u8.shape == (items, S, S, 3) and views_u8.is_contiguous() and len(sigmas) == items
    if all(sg is None for sg in sigmas):
        return views_u8
    dev = views_u8.device
    taps = torch.tensor([gaussian_taps_fixed(ksize, sg) if sg is not None else [0] * ksize for sg in sigmas], dtype=torch.int32, device=dev)
    apply = torch.tensor([int(sg is not None) for sg in sigmas], dtype=torch.int32, device=dev)
    nbytes = lib.passl_b200_gaussian_blur_workspace_bytes(items, S)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.passl_b200_gaussian_blur_u8(views_u8.data_ptr(), taps.data_ptr(), apply.data_ptr(), ws.data_ptr(), nbytes, items, S,
                                               ksize, _stream()), "gaussian_blur_u8")
    return views_u8


class ViewRecipe:
    """The per-view part of the recipe: RandomApply(ColorJitter) -> RandomGrayscale -> RandomApply(GaussianBlur) -> RandomHorizontalFlip
    -> Transpose -> NormalizeImage (the `view_trans1` / `view_trans2` lists of the reference YAMLs)."""

    def __init__(self, jitter_p=0.0, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1, gray_p=0.2, blur_p=0.0,
                 blur_sigma=(0.1, 2.0), blur_ksize=23, flip_p=0.5, norm_scale=1.0 / 255.0, mean=(0.485, 0.456, 0.406),
                 std=(0.229, 0.224, 0.225)):
        self.jitter_p, self.jitter = jitter_p, dict(brightness=brightness, contrast=contrast, saturation=saturation, hue=hue)
        self.gray_p, self.blur_p, self.blur_sigma, self.blur_ksize, self.flip_p = gray_p, blur_p, tuple(blur_sigma), blur_ksize, flip_p
        self.norm_scale, self.mean, self.std = norm_scale, tuple(mean), tuple(std)

    def draw(self, rng, np_rng):
        """-> (jitter plan, gray, blur sigma or None, flip), drawn in pipeline order.  RandomApply skips when `p < random.random()`
        (transforms.py:138-143); the blur sigma comes from numpy's generator (transforms.py:182)."""
        plan, sigma = [], None
        if self.jitter_p and not (self.jitter_p < rng.random()):
            plan = color_jitter_plan(rng=rng, **self.jitter)
        gray = rng.random() < self.gray_p
        if self.blur_p and not (self.blur_p < rng.random()):
            sigma = float(np_rng.uniform(self.blur_sigma[0], self.blur_sigma[1]))
        flip = rng.random() < self.flip_p
        return plan, gray, sigma, flip


class TwoViewInputStage:
    """`(view_1, view_2) = stage(images)`: both fp32 [N, 3, size, size] on the device, ready for MoCo / SimCLR `train_iter`."""

    def __init__(self, size=224, scale=(0.08, 1.0), ratio=(3. / 4., 4. / 3.), interpolation="bilinear", rng=random, np_rng=None,
                 view1=None, view2=None, **view_kwargs):
        """SimCLR recipe (configs/simclr/simclr_r50_IM.yaml:35-83): scale=(0.1, 1), interpolation='bicubic', jitter_p=0.8, gray_p=0.2,
        blur_p=0.5, flip_p=0.5; MoCo v2 (configs/moco/moco_v2_r50.yaml:34-80): scale=(0.2, 1), bilinear, same view lists.  Keyword
        arguments of ViewRecipe given here apply to both views; `view1` / `view2` take ready ViewRecipe objects when the two lists
        differ.  `rng` plays Python's `random` module, `np_rng` numpy's."""
        import numpy as np
        self.size, self.scale, self.ratio, self.interpolation = size, tuple(scale), tuple(ratio), interpolation
        self.views = (view1 or ViewRecipe(**view_kwargs), view2 or ViewRecipe(**view_kwargs))
        self.rng, self.np_rng = rng, (np_rng if np_rng is not None else np.random)
        self.last_sigmas = []

    def draw(self, batch):
        """Host-side random decisions for one batch -> (item_img, item_box, gray, flip, plans); views of sample n are items n and
        N + n.  Per sample: box of view 1, box of view 2 (the dataset calls the crop transform twice first, imagenet.py:57-58), then
        the decisions of view 1, then those of view 2."""
        N = len(batch)
        box1, box2, d1, d2 = [], [], [], []
        for n in range(N):
            box1.append(random_resized_crop_params(batch.widths[n], batch.heights[n], self.scale, self.ratio, self.rng))
            box2.append(random_resized_crop_params(batch.widths[n], batch.heights[n], self.scale, self.ratio, self.rng))
            d1.append(self.views[0].draw(self.rng, self.np_rng))
            d2.append(self.views[1].draw(self.rng, self.np_rng))
        both = d1 + d2
        self.last_sigmas = [d[2] for d in both]
        return list(range(N)) * 2, box1 + box2, [d[1] for d in both], [d[3] for d in both], [d[0] for d in both]

    def __call__(self, images):
        batch = images if isinstance(images, ImageBatch) else ImageBatch(images)
        item_img, item_box, gray, flip, plans = self.draw(batch)
        N = len(batch)
        u8 = resized_crop_u8(batch, item_img, item_box, self.size, self.interpolation)
        u8 = color_jitter_u8(u8, plans)
        # grayscale sits between jitter and blur in the recipe and does not commute with the blur's rounding: a view that is both
        # grey and blurred gets its grayscale here, every other grey view in the finalize kernel
        if any(sg is not None for sg in self.last_sigmas):
            pre_gray = [g and sg is not None for g, sg in zip(gray, self.last_sigmas)]
            if any(pre_gray):
                u8 = grayscale_u8(u8, pre_gray)
                gray = [g and not p for g, p in zip(gray, pre_gray)]
            ksizes = {v.blur_ksize for v in self.views}
            assert len(ksizes) == 1, "both views must use the same blur kernel size"
            u8 = gaussian_blur_u8(u8, self.last_sigmas, ksizes.pop())
        outs = []
        for v, recipe in enumerate(self.views):                          # the two view lists may normalise differently
            sl = slice(v * N, (v + 1) * N)
            outs.append(views_finalize(u8[sl], gray[sl], flip[sl], recipe.norm_scale, recipe.mean, recipe.std))
        return outs[0], outs[1]


def _fraction(text):
    """NormalizeImage's `scale: 1.0/255.0` is a string the reference eval()s (transforms.py:459); only `a` or `a/b` is accepted."""
    if isinstance(text, (int, float)):
        return float(text)
    parts = str(text).split("/")
    if len(parts) > 2:
        raise ValueError("cannot read scale %r" % (text,))
    return float(parts[0]) / (float(parts[1]) if len(parts) == 2 else 1.0)


def _view_recipe_from_list(items):
    """One `view_trans*` list -> ViewRecipe.  The list must be the supported pipeline, in its order; anything else is refused."""
    kw, stage = {}, 0
    order = ["jitter", "gray", "blur", "flip", "transpose", "normalize"]

    def at(step):
        nonlocal stage
        i = order.index(step)
        if i < stage:
            raise NotImplementedError("transform order not supported: %s after %s" % (step, order[stage - 1]))
        stage = i + 1
    for t in items:
        t = dict(t)
        name = t.pop("name")
        if name == "RandomApply":
            inner = [dict(x) for x in t["transforms"]]
            if len(inner) != 1:
                raise NotImplementedError("RandomApply with %d transforms" % len(inner))
            iname = inner[0].pop("name")
            if iname == "ColorJitter":
                at("jitter")
                kw.update(jitter_p=t.get("p", 0.5), **{k: inner[0].get(k, 0) for k in ("brightness", "contrast", "saturation", "hue")})
            elif iname == "GaussianBlur":
                at("blur")
                if inner[0].get("_PIL", False):
                    raise NotImplementedError("GaussianBlur(_PIL=True): only the default cv2 path is built")
                kw.update(blur_p=t.get("p", 0.5), blur_sigma=tuple(inner[0].get("sigma", (0.1, 2.0))))
            else:
                raise NotImplementedError("RandomApply(%s)" % iname)
        elif name == "RandomGrayscale":
            at("gray")
            kw["gray_p"] = t.get("p", 0.1)
        elif name == "RandomHorizontalFlip":
            at("flip")
            kw["flip_p"] = t.get("prob", 0.5)
        elif name == "Transpose":
            at("transpose")
        elif name == "NormalizeImage":
            at("normalize")
            kw.update(norm_scale=_fraction(t.get("scale", 1.0)), mean=tuple(t.get("mean", (0.0,) * 3)), std=tuple(t.get("std", (1.0,) * 3)))
        else:
            raise NotImplementedError("transform %s is not built in the device input stage" % name)
    kw.setdefault("gray_p", 0.0)
    kw.setdefault("flip_p", 0.0)
    return ViewRecipe(**kw)


def build_input_stage(dataset_cfg, rng=random, np_rng=None):
    """The `dataloader.train.dataset` section of a two-view YAML (configs/moco/moco_v2_r50.yaml:29-80, configs/simclr/
    simclr_r50_IM.yaml:29-83: `transforms` = [RandomResizedCrop], `view_trans1`, `view_trans2`) -> TwoViewInputStage."""
    crop = [dict(t) for t in dataset_cfg["transforms"]]
    if len(crop) != 1 or crop[0].get("name") != "RandomResizedCrop":
        raise NotImplementedError("`transforms` must be a single RandomResizedCrop, got %s" % [t.get("name") for t in crop])
    c = crop[0]
    size = c["size"]
    if isinstance(size, (list, tuple)):
        if len(set(size)) != 1:
            raise NotImplementedError("non-square output size %r" % (size,))
        size = size[0]
    return TwoViewInputStage(size=int(size), scale=tuple(c.get("scale", (0.08, 1.0))), ratio=tuple(c.get("ratio", (3. / 4., 4. / 3.))),
                             interpolation=c.get("interpolation", "bilinear"), rng=rng, np_rng=np_rng,
                             view1=_view_recipe_from_list(dataset_cfg["view_trans1"]),
                             view2=_view_recipe_from_list(dataset_cfg["view_trans2"]))


class SyntheticDecodedImages:
    """Stand-in for the decode step: per iteration a list of `batch_size` random uint8 HWC images of mixed sizes, created once on
    the device (there is no ImageNet and no JPEG decoder in this environment; sizes follow typical ImageNet aspect ratios)."""

    SHAPES = ((375, 500), (500, 375), (333, 500), (500, 333), (256, 256), (480, 640))

    def __init__(self, batch_size, iters, device, seed=1234):
        self.iters = iters
        g = torch.Generator(device=device).manual_seed(seed)
        self.images = [torch.randint(0, 256, self.SHAPES[i % len(self.SHAPES)] + (3,), dtype=torch.uint8, device=device, generator=g)
                       for i in range(batch_size)]

    def __len__(self):
        return self.iters

    def __iter__(self):
        for _ in range(self.iters):
            yield self.images


class DeviceAugmentedTwoViews:
    """Loader for the Trainer: decoded uint8 images in, `(view_1, view_2)` out, every iteration through the device input stage —
    the place of `ImageNet.__getitem__` + the DataLoader workers (passl_v110/datasets/imagenet.py:46-63)."""

    def __init__(self, source, stage):
        self.source, self.stage = source, stage
        self._packed = {}

    def __len__(self):
        return len(self.source)

    def __iter__(self):
        for images in self.source:
            key = id(images)                                         # a source that re-yields the same list is packed once
            batch = self._packed.get(key)
            if batch is None:
                batch = ImageBatch(images)
                self._packed = {key: batch}
            yield self.stage(batch)
