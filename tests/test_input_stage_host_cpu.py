"""The per-element bodies of the CUDA input-stage kernels (passl_b200/csrc/input_stage_core.h) compiled as host C++ and run here
without a GPU: ragged source images, several views per image, bilinear and bicubic, against Pillow itself (bit-exact uint8) and the
oracle's float stage (bit-exact float32).  The CUDA kernels add only the thread-index decomposition on top of this source; the
`-m gpu` test (tests/test_zz_input_stage_gpu.py) runs the same cases through the C ABI."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle.input_stage as O

HERE = os.path.dirname(os.path.abspath(__file__))
Image = pytest.importorskip("PIL.Image")


def make_cases(seed=0):
    """-> dict(src, src_off, src_h, src_w, item_img, item_box, images): three ragged images, nine views incl. the edge boxes."""
    rng = np.random.RandomState(seed)
    shapes = [(375, 500), (64, 48), (300, 224)]
    images = [rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in shapes]
    images[0][:180] = (images[0][:180] // 128) * 255                               # hard edges (bicubic over/undershoot -> clip)
    off = np.cumsum([0] + [im.size for im in images[:-1]]).astype(np.int64)
    items = [(0, (10, 20, 200, 300)), (0, (0, 0, 375, 500)), (0, (300, 400, 75, 100)), (0, (100, 100, 1, 1)),
             (1, (0, 0, 64, 48)), (1, (5, 7, 30, 20)), (2, (0, 0, 300, 224)), (2, (38, 0, 224, 224)), (2, (299, 223, 1, 1))]
    return dict(src=np.concatenate([im.reshape(-1) for im in images]), src_off=off,
                src_h=np.array([s[0] for s in shapes], dtype=np.int32), src_w=np.array([s[1] for s in shapes], dtype=np.int32),
                item_img=np.array([i for i, _ in items], dtype=np.int32), item_box=np.array([b for _, b in items], dtype=np.int32),
                images=images)


def kmax_for(max_crop, S, bicubic):
    return int(np.ceil((2.0 if bicubic else 1.0) * max(1.0, max_crop / S))) * 2 + 1


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("host_istage") / "libhost_istage.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++14", "-o", out,
                           os.path.join(HERE, "host_input_stage.cpp")])
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("bicubic", [0, 1])
@pytest.mark.parametrize("S", [224, 32])
def test_host_build_of_the_kernel_bodies_is_pillow_exact(hostlib, bicubic, S):
    c = make_cases()
    n = len(c["item_img"])
    max_h, max_side = int(c["item_box"][:, 2].max()), int(c["item_box"][:, 2:].max())
    dst = np.full((n, S, S, 3), 77, dtype=np.uint8)
    st = hostlib.host_resized_crop_u8(_p(c["src"]), _p(c["src_off"]), _p(c["src_h"]), _p(c["src_w"]), _p(c["item_img"]),
                                      _p(c["item_box"]), _p(dst), n, S, max_h, kmax_for(max_side, S, bicubic), bicubic)
    assert st == 0
    name = "bicubic" if bicubic else "bilinear"
    for m in range(n):
        i, j, h, w = (int(v) for v in c["item_box"][m])
        pil = Image.fromarray(c["images"][c["item_img"][m]]).crop((j, i, j + w, i + h)).resize((S, S), getattr(Image, name.upper()))
        assert np.array_equal(dst[m], np.asarray(pil)), (m, name)
        assert np.array_equal(dst[m], O.resized_crop_u8(c["images"][c["item_img"][m]], i, j, h, w, S, name))
    # float stage: grayscale / flip / CHW / normalise
    gray = np.array([0, 1, 0, 1, 0, 0, 1, 0, 1], dtype=np.int32)
    flip = np.array([0, 0, 1, 1, 0, 1, 0, 1, 0], dtype=np.int32)
    mean, std = np.float32([0.485, 0.456, 0.406]), np.float32([0.229, 0.224, 0.225])
    out = np.zeros((n, 3, S, S), dtype=np.float32)
    hostlib.host_views_finalize_f32(_p(dst), _p(gray), _p(flip), _p(out), n, S, ctypes.c_double(1.0 / 255.0), _p(mean), _p(std))
    for m in range(n):
        img = dst[m]
        if gray[m]:
            img = O.grayscale3_u8(img)
        if flip[m]:
            img = O.hflip_u8(img)
        assert np.array_equal(out[m], O.transpose_normalize(img)), m


def test_status_word_reports_bad_boxes_and_short_kmax(hostlib):
    c = make_cases()
    n, S = len(c["item_img"]), 16
    dst = np.zeros((n, S, S, 3), dtype=np.uint8)
    box = c["item_box"].copy()
    box[2] = (300, 400, 76, 100)                                                     # one row past the image
    st = hostlib.host_resized_crop_u8(_p(c["src"]), _p(c["src_off"]), _p(c["src_h"]), _p(c["src_w"]), _p(c["item_img"]), _p(box),
                                      _p(dst), n, S, 375, kmax_for(500, S, 0), 0)
    assert st == 1 and dst[2].max() == 0 and dst[0].max() > 0
    st = hostlib.host_resized_crop_u8(_p(c["src"]), _p(c["src_off"]), _p(c["src_h"]), _p(c["src_w"]), _p(c["item_img"]),
                                      _p(c["item_box"]), _p(dst), n, S, 375, 5, 0)
    assert st == 2 and dst[1].max() == 0 and dst[3].max() > 0                        # 500 -> 16 needs 65 taps; the 1x1 crop needs 3


def test_product_sampler_and_draw_order():
    """passl_b200.data.random_resized_crop_params equals the reference's get_params (same `random` calls in the same order), and
    TwoViewInputStage.draw consumes the stream as documented: box 1, box 2, gray 1, flip 1, gray 2, flip 2 per sample."""
    import random
    from passl_b200.data import ImageBatch, TwoViewInputStage, random_resized_crop_params
    G = np.load(os.path.join(HERE, "golden", "reference_crop_params.npz"))
    for tag in "abcdef":
        W, H, s0, s1, r0, r1, seed = G["args_" + tag]
        random.seed(int(seed))
        got = [random_resized_crop_params(int(W), int(H), (s0, s1), (r0, r1), random) for _ in range(64)]
        assert np.array_equal(np.array(got), G["boxes_" + tag]), tag
    images = [np.zeros((h, w, 3), dtype=np.uint8) for h, w in [(240, 320), (100, 50)]]
    batch = ImageBatch(images, device="cpu")
    assert batch.src_off.tolist() == [0, 240 * 320 * 3] and batch.data.numel() == 240 * 320 * 3 + 100 * 50 * 3
    stage = TwoViewInputStage(size=64, jitter_p=0.8, blur_p=0.7, rng=random.Random(3), np_rng=np.random.RandomState(8))
    item_img, item_box, gray, flip, plans = stage.draw(batch)
    r, nr = random.Random(3), np.random.RandomState(8)

    def view():                                          # RandomApply(ColorJitter) -> RandomGrayscale -> RandomHorizontalFlip
        plan = []
        if not (0.8 < r.random()):
            slots = [("brightness", 0.4), ("contrast", 0.4), ("saturation", 0.4), ("hue", 0.1)]
            r.shuffle(slots)
            for name, v in slots:
                lo, hi = (-v, v) if name == "hue" else (max(0.0, 1 - v), 1 + v)
                plan.append(({"brightness": 1, "contrast": 2, "saturation": 3, "hue": 4}[name], r.uniform(lo, hi)))
        g = r.random() < 0.2
        sigma = float(nr.uniform(0.1, 2.0)) if not (0.7 < r.random()) else None       # RandomApply(GaussianBlur) + np.random sigma
        return plan, g, r.random() < 0.5, sigma
    want = {}
    for n, (h, w) in enumerate([(240, 320), (100, 50)]):
        b1, b2 = random_resized_crop_params(w, h, rng=r), random_resized_crop_params(w, h, rng=r)
        want[n], want[2 + n] = (b1,) + view(), (b2,) + view()
    assert item_img == [0, 1, 0, 1]
    for m in range(4):
        assert (tuple(item_box[m]), plans[m], gray[m], flip[m], stage.last_sigmas[m]) == want[m], m
    assert any(sg is not None for sg in stage.last_sigmas) and any(sg is None for sg in stage.last_sigmas)
    assert any(len(p) == 4 for p in plans) and all(sorted(op for op, _ in p) == [1, 2, 3, 4] for p in plans if p)
    with pytest.raises(ValueError):
        ImageBatch([np.zeros((4, 4), dtype=np.uint8)], device="cpu")


def hue_shift(hue_factor):
    """np.uint8(hue_factor * 255) of paddle's adjust_hue: truncation toward zero, then wrap"""
    return int(hue_factor * 255) & 255


def jitter_cases(seed=3, items=6, S=24):
    """random views + a 4-op plan per view (shuffled order, some ops absent) -> (img u8 [items,S,S,3], ops, factors, plan)"""
    import random
    rng, r = np.random.RandomState(seed), random.Random(seed)
    img = rng.randint(0, 256, size=(items, S, S, 3)).astype(np.uint8)
    img[1, : S // 2] //= 5
    img[2] = 200                                                                       # flat view: grey -> hue / saturation no-ops
    ops, factors, plan = np.zeros((items, 4), dtype=np.int32), np.zeros((items, 4), dtype=np.float32), []
    for m in range(items):
        order = [1, 2, 3, 4]
        r.shuffle(order)
        if m == 4:
            order = [2, 0, 2, 0]                                                       # contrast twice: the mean must be re-taken
        row = []
        for pos, op in enumerate(order):
            f = r.uniform(-0.1, 0.1) if op == 4 else r.uniform(0.6, 1.4)
            if m == 5 and pos == 0:
                f = 1.9 if op != 4 else 0.5                                            # extrapolating blend -> clipping branch
            ops[m, pos] = op
            factors[m, pos] = hue_shift(f) if op == 4 else f
            row.append((op, f))
        plan.append(row)
    return img, ops, factors, plan


def oracle_jitter(view, row):
    for op, f in row:
        if op == 1:
            view = O.adjust_brightness(view, float(np.float32(f)))
        elif op == 2:
            view = O.adjust_contrast(view, float(np.float32(f)))
        elif op == 3:
            view = O.adjust_saturation(view, float(np.float32(f)))
        elif op == 4:
            view = O.adjust_hue(view, f)
    return view


def test_host_build_of_colour_jitter_bodies(hostlib):
    v = np.arange(1 << 24, dtype=np.uint32)
    cube = np.ascontiguousarray(np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], -1).astype(np.uint8))
    out = np.empty_like(cube)
    hostlib.host_rgb_to_hsv(_p(cube), _p(out), ctypes.c_longlong(cube.shape[0]))
    assert np.array_equal(out.reshape(4096, 4096, 3), np.asarray(Image.fromarray(cube.reshape(4096, 4096, 3), "RGB").convert("HSV")))
    hostlib.host_hsv_to_rgb(_p(cube), _p(out), ctypes.c_longlong(cube.shape[0]))
    assert np.array_equal(out.reshape(4096, 4096, 3), np.asarray(Image.fromarray(cube.reshape(4096, 4096, 3), "HSV").convert("RGB")))
    img, ops, factors, plan = jitter_cases()
    got = img.copy()
    hostlib.host_color_jitter_u8(_p(got), _p(ops), _p(factors), img.shape[0], img.shape[1])
    for m in range(img.shape[0]):
        assert np.array_equal(got[m], oracle_jitter(img[m], plan[m])), (m, plan[m])
    assert not np.array_equal(got[0], img[0]) and np.array_equal(got[2][..., 0], got[2][..., 1])
    # op 5: the in-place grayscale used for views that are blurred afterwards
    g_ops = np.zeros((img.shape[0], 4), dtype=np.int32)
    g_ops[[0, 3], 0] = 5
    got = img.copy()
    hostlib.host_color_jitter_u8(_p(got), _p(g_ops), _p(np.zeros((img.shape[0], 4), dtype=np.float32)), img.shape[0], img.shape[1])
    for m in range(img.shape[0]):
        assert np.array_equal(got[m], O.grayscale3_u8(img[m]) if m in (0, 3) else img[m]), m


def test_host_build_of_gaussian_blur_bodies_is_opencv_exact(hostlib):
    cv2 = pytest.importorskip("cv2")
    import random
    from passl_b200.data import gaussian_taps_fixed
    rng, r = np.random.RandomState(2), random.Random(2)
    for S in (24, 64):
        items = 5
        img = rng.randint(0, 256, size=(items, S, S, 3)).astype(np.uint8)
        img[1, : S // 2] = (img[1, : S // 2] // 128) * 255
        sigmas = [0.1, 2.0, None, r.uniform(0.1, 2.0), r.uniform(0.1, 2.0)]
        for sg in sigmas:
            if sg is not None:
                assert gaussian_taps_fixed(23, sg) == [int(v) for v in O.gaussian_taps_fixed(23, sg)]
        taps = np.array([gaussian_taps_fixed(23, sg) if sg is not None else [0] * 23 for sg in sigmas], dtype=np.int32)
        apply = np.array([int(sg is not None) for sg in sigmas], dtype=np.int32)
        got = img.copy()
        hostlib.host_gaussian_blur_u8(_p(got), _p(taps), _p(apply), items, S, 23)
        for m, sg in enumerate(sigmas):
            want = img[m] if sg is None else cv2.GaussianBlur(img[m], (23, 23), sg)
            assert np.array_equal(got[m], want), (S, m, sg)
    # a view narrower than the kernel radius: repeated reflection like cv::borderInterpolate
    small = rng.randint(0, 256, size=(1, 7, 7, 3)).astype(np.uint8)
    taps = np.array([gaussian_taps_fixed(23, 1.7)], dtype=np.int32)
    got = small.copy()
    hostlib.host_gaussian_blur_u8(_p(got), _p(taps), _p(np.array([1], dtype=np.int32)), 1, 7, 23)
    assert np.array_equal(got[0], cv2.GaussianBlur(small[0], (23, 23), 1.7))
    with pytest.raises(ValueError):
        gaussian_taps_fixed(22, 1.0)


def test_build_input_stage_from_the_reference_yaml_sections():
    """`dataloader.train.dataset` of the MoCo v2 / SimCLR YAMLs (the reference's own transform lists) -> the device stage."""
    from passl_b200.data import build_input_stage
    from passl_b200.utils.config import get_config
    root = os.path.dirname(HERE)
    for f, scale, interp in (("configs/moco/moco_v2_r50.yaml", (0.2, 1.0), "bilinear"), ("configs/simclr/simclr_r50_IM.yaml", (0.1, 1.0), "bicubic")):
        st = build_input_stage(get_config(os.path.join(root, f), []).dataloader.train.dataset)
        assert (st.size, st.scale, st.ratio, st.interpolation) == (224, scale, (3. / 4., 4. / 3.), interp)
        for v in st.views:
            assert (v.jitter_p, v.gray_p, v.blur_p, v.blur_sigma, v.blur_ksize, v.flip_p) == (0.8, 0.2, 0.5, (0.1, 2.0), 23, 0.5)
            assert v.jitter == dict(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1)
            assert v.norm_scale == 1.0 / 255.0 and v.mean == (0.485, 0.456, 0.406) and v.std == (0.229, 0.224, 0.225)
    crop = [dict(name="RandomResizedCrop", size=96)]
    plain = [dict(name="RandomHorizontalFlip"), dict(name="Transpose"), dict(name="NormalizeImage", scale="1.0/255.0", mean=[0.5] * 3, std=[0.5] * 3)]
    st = build_input_stage(dict(transforms=crop, view_trans1=plain, view_trans2=[dict(name="RandomGrayscale")] + plain))
    assert st.views[0].gray_p == 0.0 and st.views[1].gray_p == 0.1 and st.views[0].jitter_p == 0.0 and st.scale == (0.08, 1.0)
    with pytest.raises(NotImplementedError):                                           # an op the stage does not have
        build_input_stage(dict(transforms=crop, view_trans1=[dict(name="Solarization")], view_trans2=plain))
    with pytest.raises(NotImplementedError):                                           # supported ops in another order
        build_input_stage(dict(transforms=crop, view_trans1=[dict(name="RandomHorizontalFlip"), dict(name="RandomGrayscale")], view_trans2=plain))
    with pytest.raises(NotImplementedError):
        build_input_stage(dict(transforms=[dict(name="Resize", size=224)], view_trans1=plain, view_trans2=plain))


@pytest.fixture()
def stage_on_host(hostlib, monkeypatch):
    """passl_b200.data's wrappers bound to the host build: same ABI names / ctypes signatures, host memory, no stream."""
    from passl_b200 import _lib
    from passl_b200.data import input_stage as IS
    for name in ("passl_b200_resample_kmax", "passl_b200_resized_crop_workspace_bytes", "passl_b200_resized_crop_u8",
                 "passl_b200_views_finalize_f32", "passl_b200_color_jitter_u8", "passl_b200_gaussian_blur_workspace_bytes",
                 "passl_b200_gaussian_blur_u8"):
        fn = getattr(hostlib, name)
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]
    monkeypatch.setattr(IS._lib, "load", lambda: hostlib)
    monkeypatch.setattr(IS, "_stream", lambda: 0)
    monkeypatch.setattr(IS, "_need_cuda", lambda t: None)
    return IS


def test_python_wrappers_and_stage_orchestration_on_the_host_build(stage_on_host):
    """The GPU tests of tests/test_zz_input_stage_gpu.py, with the library swapped for its host build: argument marshalling of every
    wrapper, the status word, and TwoViewInputStage.__call__ (jitter -> grayscale-before-blur -> blur -> finalize per view list)."""
    import random
    import torch
    IS = stage_on_host
    c = make_cases()
    batch = IS.ImageBatch(c["images"], device="cpu")
    boxes = [tuple(int(v) for v in b) for b in c["item_box"]]
    idx = [int(i) for i in c["item_img"]]
    for interp in ("bilinear", "bicubic"):
        u8 = IS.resized_crop_u8(batch, idx, boxes, 32, interp)
        for m, (n, (i, j, h, w)) in enumerate(zip(idx, boxes)):
            assert np.array_equal(u8[m].numpy(), O.resized_crop_u8(c["images"][n], i, j, h, w, 32, interp)), (m, interp)
    with pytest.raises(ValueError):
        IS.resized_crop_u8(batch, [0], [(300, 400, 76, 100)], 32)
    batch.heights[0] += 1                                                               # lie to the host check: device status word
    with pytest.raises(IS._lib.PasslB200Error):
        IS.resized_crop_u8(batch, [0, 1], [(300, 400, 76, 100), (0, 0, 64, 48)], 32)
    batch.heights[0] -= 1
    img, ops, factors, plan = jitter_cases()
    got = IS.color_jitter_u8(torch.from_numpy(img.copy()), [[(op, f) for op, f in row if op] for row in plan]).numpy()
    for m in range(img.shape[0]):
        assert np.array_equal(got[m], oracle_jitter(img[m], [(op, f) for op, f in plan[m] if op])), m
    sig = [0.1, None, 1.37, 2.0, None, 0.5]
    got = IS.gaussian_blur_u8(torch.from_numpy(img.copy()), sig).numpy()
    for m, sg in enumerate(sig):
        assert np.array_equal(got[m], img[m] if sg is None else O.gaussian_blur_u8(img[m], 23, sg)), m
    # the whole two-view stage, both view lists different (view 2 never blurs, other normalisation)
    rng = np.random.RandomState(5)
    images = [rng.randint(0, 256, size=(int(h), int(w), 3)).astype(np.uint8) for h, w in [(240, 320), (333, 250), (128, 128), (96, 400)]]
    v1 = dict(jitter_p=0.6, gray_p=0.5, blur_p=0.6)
    v2 = dict(jitter_p=0.6, gray_p=0.5, blur_p=0.0, mean=(0.5, 0.5, 0.5), std=(0.25, 0.25, 0.25))
    mk = lambda: IS.TwoViewInputStage(size=48, scale=(0.2, 1.0), interpolation="bicubic", rng=random.Random(11),   # noqa: E731
                                      np_rng=np.random.RandomState(4), view1=IS.ViewRecipe(**v1), view2=IS.ViewRecipe(**v2))
    stage, twin = mk(), mk()
    b2 = IS.ImageBatch(images, device="cpu")
    out1, out2 = stage(b2)
    item_img, item_box, gray, flip, plans = twin.draw(b2)
    sig = twin.last_sigmas
    assert any(s_ is not None and g for s_, g in zip(sig, gray)) and any(s_ is None and g for s_, g in zip(sig, gray)) and any(plans)
    both = torch.cat([out1, out2]).numpy()
    for m, (n, (i, j, h, w)) in enumerate(zip(item_img, item_box)):
        x = oracle_jitter(O.resized_crop_u8(images[n], i, j, h, w, 48, "bicubic"), plans[m])
        if gray[m]:
            x = O.grayscale3_u8(x)
        if sig[m] is not None:
            x = O.gaussian_blur_u8(x, 23, sig[m])
        if flip[m]:
            x = O.hflip_u8(x)
        kw = {} if m < 4 else dict(mean=v2["mean"], std=v2["std"])
        assert np.array_equal(both[m], O.transpose_normalize(x, **kw)), m
