"""Generate the golden vectors under tests/golden/ by running the REFERENCE'S OWN Python source (read in place from
/root/reference, never copied) over the torch-backed paddle shim (paddle_shim.py).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests consume the committed .npz files.  Each file stores the seeded inputs and the
outputs the reference code produced, in float64.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
import paddle_shim  # noqa: E402


def setup():
    paddle_shim.install()
    paddle_shim.install_finder()
    # packages of the v110 tree: expose directories without running their __init__ (which imports the whole zoo)
    for name, sub in [("passl_v110", ""), ("passl_v110.utils", "utils"), ("passl_v110.modeling", "modeling"),
                      ("passl_v110.modeling.heads", "modeling/heads"), ("passl_v110.modeling.architectures", "modeling/architectures"),
                      ("passl_v110.modeling.backbones", "modeling/backbones"), ("passl_v110.modeling.necks", "modeling/necks"),
                      ("passl_v110.modules", "modules")]:
        paddle_shim.fake_package(name, os.path.join(REF, "passl_v110", sub))
    fm = importlib.import_module("passl_v110.modules.freeze")
    sys.modules["passl_v110.modules"].freeze_batchnorm_statictis = fm.freeze_batchnorm_statictis
    for pkg, fn in [("backbones", "build_backbone"), ("necks", "build_neck"), ("heads", "build_head")]:
        b = importlib.import_module("passl_v110.modeling.%s.builder" % pkg)
        setattr(sys.modules["passl_v110.modeling.%s" % pkg], fn, getattr(b, fn))
    for name, sub in [("passl", ""), ("passl.models", "models"), ("passl.models.utils", "models/utils"), ("passl.nn", "nn")]:
        paddle_shim.fake_package(name, os.path.join(REF, "passl", sub))


def t(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float64))


def gen_contrastive_head(out):
    mod = importlib.import_module("passl_v110.modeling.heads.contrastive_head")
    rng = np.random.RandomState(1234)
    for tag, (N, K, T) in {"a": (16, 4096, 0.2), "b": (37, 1000, 0.07)}.items():
        pos = rng.randn(N, 1) * 0.3 + 0.5
        neg = rng.randn(N, K) * 0.3
        head = mod.ContrastiveHead(temperature=T)
        o = head(t(pos), t(neg))
        out["contrastive_%s_pos" % tag] = pos
        out["contrastive_%s_neg" % tag] = neg
        out["contrastive_%s_T" % tag] = np.float64(T)
        out["contrastive_%s_loss" % tag] = o["loss"].numpy()
        out["contrastive_%s_acc1" % tag] = o["acc1"].numpy()
        out["contrastive_%s_acc5" % tag] = o["acc5"].numpy()


def gen_simclr_head(out):
    mod = importlib.import_module("passl_v110.modeling.heads.simclr_contrastive_head")
    rng = np.random.RandomState(4321)
    for tag, (n, d, T) in {"a": (24, 32, 0.1), "b": (64, 128, 0.5)}.items():
        h1 = rng.randn(n, d)
        h2 = 0.5 * h1 + 0.5 * rng.randn(n, d)
        h1 /= np.linalg.norm(h1, axis=1, keepdims=True)
        h2 /= np.linalg.norm(h2, axis=1, keepdims=True)
        head = mod.SimCLRContrastiveHead(temperature=T)
        o = head(t(h1), t(h2))
        out["simclr_%s_h1" % tag] = h1
        out["simclr_%s_h2" % tag] = h2
        out["simclr_%s_T" % tag] = np.float64(T)
        out["simclr_%s_loss" % tag] = o["loss"].numpy()
        out["simclr_%s_acc1" % tag] = np.asarray(o["acc1"].numpy())


def gen_clip_head(out):
    mod = importlib.import_module("passl_v110.modeling.heads.clip_head")
    rng = np.random.RandomState(99)
    n, d = 20, 64
    img = rng.randn(n, d)
    txt = 0.7 * img + 0.3 * rng.randn(n, d)
    ls = np.log(1 / 0.07)
    i = img / np.linalg.norm(img, axis=-1, keepdims=True)
    x = txt / np.linalg.norm(txt, axis=-1, keepdims=True)
    s = np.exp(ls)
    img_logits, text_logits = (s * i) @ x.T, (s * x) @ i.T           # clip.py:331-335 (plain matmuls)
    head = mod.CLIPHead()
    labels = torch.arange(n)
    o = head(t(img_logits), t(text_logits), labels, labels)
    out["clip_img"], out["clip_txt"], out["clip_logit_scale"] = img, txt, np.float64(ls)
    out["clip_img_loss"], out["clip_text_loss"], out["clip_loss"] = o["img_loss"].numpy(), o["text_loss"].numpy(), o["loss"].numpy()


def gen_moco_queue(out):
    mod = importlib.import_module("passl_v110.modeling.architectures.moco")
    rng = np.random.RandomState(7)
    D, K, B = 16, 64, 8
    m = mod.MoCo.__new__(mod.MoCo)
    torch.nn.Module.__init__(m)
    m.K = K
    m.register_buffer("queue", t(rng.randn(D, K)))
    m.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.int64))
    out["queue_init"] = m.queue.numpy().copy()
    keys_all, ptrs, snaps = [], [], []
    for step in range(10):                      # wraps around once (10*8 = 80 > 64)
        keys = rng.randn(B, D)
        m._dequeue_and_enqueue(t(keys))
        keys_all.append(keys)
        ptrs.append(int(m.queue_ptr[0]))
        snaps.append(m.queue.numpy().copy())
    out["queue_keys"] = np.stack(keys_all)
    out["queue_ptrs"] = np.asarray(ptrs, dtype=np.int64)
    out["queue_final"] = snaps[-1]
    out["queue_after3"] = snaps[2]
    # momentum update formula, moco.py:82-90
    pk, pq = rng.randn(50), rng.randn(50)
    out["ema_pk"], out["ema_pq"] = pk, pq
    out["ema_out"] = (t(pk) * 0.999 + t(pq) * (1. - 0.999)).numpy()


def gen_pos_embed(out):
    mod = importlib.import_module("passl.models.utils.pos_embed")
    out["pos_embed_768_14"] = mod.get_2d_sincos_pos_embed(768, 14, cls_token=True)
    out["pos_embed_512_14"] = mod.get_2d_sincos_pos_embed(512, 14, cls_token=True)


def gen_mae(out):
    """MAE patchify / random_masking / forward_loss executed from passl/models/mae.py (unbound, on a stub self)."""
    try:
        mod = importlib.import_module("passl.models.mae")
    except Exception as e:  # the module pulls in the whole ViT stack; record why if the shim cannot import it
        out["mae_import_error"] = np.asarray(str(e))
        return
    cls = mod.MaskedAutoencoderViT
    rng = np.random.RandomState(5)

    class Stub:
        pass
    s = Stub()
    s.patch_embed = Stub()
    s.patch_embed.patch_size = (16, 16)
    N = 3
    imgs = rng.randn(N, 3, 64, 64)
    pred = rng.randn(N, 16, 768)
    noise = rng.rand(N, 16)
    s.patchify = lambda im: cls.patchify(s, im)
    target = cls.patchify(s, t(imgs))
    out["mae_imgs"], out["mae_pred"], out["mae_noise"] = imgs, pred, noise
    out["mae_patchify"] = target.numpy()
    # random_masking with the noise supplied (paddle.rand is the only RNG call)
    import paddle
    paddle.rand = lambda shape: t(noise)
    x = rng.randn(N, 16, 8)
    xm, mask, ids_restore = cls.random_masking(s, t(x), 0.75)
    out["mae_x"], out["mae_x_masked"], out["mae_mask"], out["mae_ids_restore"] = x, xm.numpy(), mask.numpy(), ids_restore.numpy()
    for npl in (False, True):
        s.norm_pix_loss = npl
        loss = cls.forward_loss(s, t(imgs), t(pred), mask)
        out["mae_loss_normpix%d" % int(npl)] = loss.numpy()


def main():
    setup()
    out = {}
    for fn in (gen_contrastive_head, gen_simclr_head, gen_clip_head, gen_moco_queue, gen_pos_embed, gen_mae):
        fn(out)
        print("generated", fn.__name__)
    np.savez_compressed(os.path.join(HERE, "reference_heads.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_heads.npz"), "with", len(out), "arrays")
    if "mae_import_error" in out:
        print("MAE import failed:", out["mae_import_error"])


if __name__ == "__main__":
    main()
