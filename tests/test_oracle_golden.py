"""Pin the CPU oracle (oracle/) against the golden vectors produced by the reference's own source
(tests/golden/make_golden.py -> tests/golden/reference_heads.npz).  CPU only."""
import os

import numpy as np
import pytest

from oracle import contrastive as OC
from oracle import mae as OM

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_heads.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_contrastive_head(tag):
    o = OC.contrastive_head(G["contrastive_%s_pos" % tag], G["contrastive_%s_neg" % tag], float(G["contrastive_%s_T" % tag]))
    np.testing.assert_allclose(o["loss"], G["contrastive_%s_loss" % tag], rtol=1e-12)
    np.testing.assert_allclose(o["acc1"], G["contrastive_%s_acc1" % tag].item(), rtol=1e-12)
    np.testing.assert_allclose(o["acc5"], G["contrastive_%s_acc5" % tag].item(), rtol=1e-12)
    assert o["labels"].dtype == np.int64 and not o["labels"].any()          # MoCo labels: all-zero int64


@pytest.mark.parametrize("tag", ["a", "b"])
def test_simclr_head(tag):
    o = OC.simclr_head(G["simclr_%s_h1" % tag], G["simclr_%s_h2" % tag], float(G["simclr_%s_T" % tag]))
    np.testing.assert_allclose(o["loss"], G["simclr_%s_loss" % tag], rtol=1e-10)
    np.testing.assert_allclose(o["acc1"], G["simclr_%s_acc1" % tag], rtol=1e-12)


def test_clip_head():
    il, tl = OC.clip_logits(G["clip_img"], G["clip_txt"], float(G["clip_logit_scale"]))
    n = il.shape[0]
    o = OC.clip_head(il, tl, np.arange(n), np.arange(n))
    np.testing.assert_allclose(o["img_loss"], G["clip_img_loss"], rtol=1e-9)
    np.testing.assert_allclose(o["text_loss"], G["clip_text_loss"], rtol=1e-9)
    np.testing.assert_allclose(o["loss"], G["clip_loss"], rtol=1e-9)


def test_queue_protocol_bit_exact():
    q, ptr = G["queue_init"].copy(), np.int64(0)
    K = q.shape[1]
    for step, keys in enumerate(G["queue_keys"]):
        q, ptr = OC.dequeue_and_enqueue(q, ptr, keys)
        assert int(ptr) == int(G["queue_ptrs"][step]) == ((step + 1) * keys.shape[0]) % K
        if step == 2:
            assert np.array_equal(q, G["queue_after3"])
    assert np.array_equal(q, G["queue_final"])
    with pytest.raises(AssertionError):
        OC.dequeue_and_enqueue(q, ptr, np.zeros((5, q.shape[0])))            # K % batch != 0 (moco.py:99)
    np.testing.assert_allclose(OC.momentum_update(G["ema_pk"], G["ema_pq"], 0.999), G["ema_out"], rtol=1e-15)


def test_pos_embed_bit_exact():
    assert np.array_equal(OM.get_2d_sincos_pos_embed(768, 14, True), G["pos_embed_768_14"])
    assert np.array_equal(OM.get_2d_sincos_pos_embed(512, 14, True), G["pos_embed_512_14"])


def test_mae_masking_and_loss():
    assert np.array_equal(OM.patchify(G["mae_imgs"]), G["mae_patchify"])
    np.testing.assert_array_equal(OM.unpatchify(OM.patchify(G["mae_imgs"])), G["mae_imgs"])
    xm, mask, ids = OM.random_masking(G["mae_x"], 0.75, G["mae_noise"])
    assert np.array_equal(ids, G["mae_ids_restore"]) and ids.dtype == np.int64
    assert np.array_equal(mask, G["mae_mask"]) and np.array_equal(xm, G["mae_x_masked"])
    N, L = mask.shape
    assert mask.sum() == N * (L - int(L * 0.25))
    for npl in (0, 1):
        got = OM.forward_loss(G["mae_imgs"], G["mae_pred"], G["mae_mask"], bool(npl))
        np.testing.assert_allclose(got, G["mae_loss_normpix%d" % npl], rtol=1e-12)


def test_known_answers():
    """SURVEY.md §8c item 4."""
    K, N, D, T = 65536, 4, 128, 0.2
    z = np.zeros((N, D))
    o = OC.contrastive_head(*OC.moco_logits(z, z, np.zeros((D, K))), T)
    np.testing.assert_allclose(o["loss"], np.log(K + 1), rtol=1e-12)
    q = np.zeros((N, D)); q[:, 0] = 1
    queue = np.zeros((D, K)); queue[1, :] = 1
    o = OC.contrastive_head(*OC.moco_logits(q, q, queue), T)
    np.testing.assert_allclose(o["loss"], np.log(1 + K * np.exp(-1 / T)), rtol=1e-12)
    # sharded (world = 4) MoCo v3 loss equals the unsharded one
    rng = np.random.RandomState(0)
    qa, ka = rng.randn(32, 16), rng.randn(32, 16)
    full, _, _ = OC.mocov3_contrastive_loss(qa, ka, 0.2, rank=0)
    parts = [OC.mocov3_contrastive_loss(qa[r * 8:(r + 1) * 8], ka, 0.2, rank=r)[0] for r in range(4)]
    np.testing.assert_allclose(np.mean(parts), full, rtol=1e-12)


def test_clip_torch_twin_matches_reference_head():
    """oracle/clip.py (the differentiable torch twin used by the GPU model tests) reproduces the reference CLIPHead golden values and
    the logit_scale clamp of clip.py:316-318."""
    import torch
    import oracle.clip as OCL
    img = torch.tensor(G["clip_img"], dtype=torch.float64)
    txt = torch.tensor(G["clip_txt"], dtype=torch.float64)
    ls = torch.tensor([float(G["clip_logit_scale"])], dtype=torch.float64)
    il, tl, ls_after = OCL.clip_forward(img, txt, ls)
    o = OCL.clip_head(il, tl)
    np.testing.assert_allclose(o["img_loss"].item(), G["clip_img_loss"], rtol=1e-9)
    np.testing.assert_allclose(o["text_loss"].item(), G["clip_text_loss"], rtol=1e-9)
    np.testing.assert_allclose(o["loss"].item(), G["clip_loss"], rtol=1e-9)
    assert ls_after.item() == ls.item()                                  # ln(1/0.07) = 2.659 is inside [-4.6, 4.6]
    assert OCL.clip_forward(img, txt, torch.tensor([7.0], dtype=torch.float64))[2].item() == 4.6
    m = OCL.build_attention_mask(5)
    assert torch.isinf(m[0, 1]) and m[1, 0] == 0 and m[3, 3] == 0          # clip.py:293-295: strictly upper triangle = -inf


def test_mocov3_contrastive_loss_vs_reference_method():
    """oracle mocov3_contrastive_loss == MoCoV3Pretrain.contrastive_loss of the reference source (tests/golden/reference_mocov3.npz,
    generated by tests/golden/make_golden_models.py calling the reference method over the shim)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_mocov3.npz"))
    for tag in ("a", "b"):
        loss, _, labels = OC.mocov3_contrastive_loss(g["q_" + tag], g["k_" + tag], float(g["T_" + tag]), rank=0)
        np.testing.assert_allclose(loss, g["loss_" + tag], rtol=1e-12)
        assert np.array_equal(labels, np.arange(g["q_" + tag].shape[0]))
