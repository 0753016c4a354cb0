"""MAE pieces and the whole MAE model on the GPU vs the oracle (oracle/mae.py pinned to the reference source through
tests/golden; oracle/vit.py = differentiable torch twin)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_heads.npz"))


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def test_random_masking_bit_exact_vs_reference_golden():
    from passl_b200 import kernels_vit as V
    noise = torch.from_numpy(G["mae_noise"]).float().cuda()
    # golden noise is float64; the kernel sorts the float32 values: compare against the oracle on the same float32 noise
    from oracle import mae as OM
    n32 = noise.cpu().numpy()
    x = G["mae_x"]
    xm, mask_ref, ids_ref = OM.random_masking(x, 0.75, n32)
    L = n32.shape[1]
    ids_shuffle, ids_restore, mask = V.mae_random_masking(noise, int(L * 0.25))
    torch.cuda.synchronize()
    assert ids_restore.dtype == torch.int64
    assert np.array_equal(ids_restore.cpu().numpy(), ids_ref)                 # bit-exact integer indices
    assert np.array_equal(mask.cpu().numpy(), mask_ref.astype(np.float32))
    assert np.array_equal(ids_restore.cpu().numpy(), G["mae_ids_restore"])    # and equal to the reference source's output
    # larger case: 196 patches, 64 samples; permutation + mask count invariants
    nz = torch.rand(64, 196, device="cuda")
    s, r, m = V.mae_random_masking(nz, 49)
    assert torch.equal(torch.sort(r, dim=1).values, torch.arange(196, device="cuda").expand(64, -1))
    assert m.sum().item() == 64 * (196 - 49)
    assert torch.equal(torch.argsort(nz, dim=1, stable=True), s)


@pytest.mark.parametrize("norm_pix", [False, True])
def test_mae_loss_matches_reference_golden(norm_pix):
    from oracle import mae as OM
    from passl_b200 import kernels_vit as V
    imgs = torch.from_numpy(G["mae_imgs"]).float().cuda()
    pred = torch.from_numpy(G["mae_pred"]).float().cuda().bfloat16()
    mask = torch.from_numpy(G["mae_mask"]).float().cuda()
    B, L = mask.shape
    Hp = int(L ** .5)
    ms = float(mask.sum().item())
    loss = V.mae_loss_fwd(pred.reshape(B * L, -1), imgs, mask, B, Hp, 16, L, 0, norm_pix, ms)
    ref_bf16 = OM.forward_loss(G["mae_imgs"].astype(np.float32).astype(np.float64), pred.float().cpu().numpy().astype(np.float64),
                               G["mae_mask"], norm_pix)
    assert abs(loss.item() - ref_bf16) < 1e-4 * abs(ref_bf16)
    assert abs(loss.item() - float(G["mae_loss_normpix%d" % int(norm_pix)])) < 1e-2 * abs(ref_bf16)   # vs reference output (bf16 pred)
    dl = torch.tensor([1.0], device="cuda")
    dpred = V.mae_loss_bwd(pred.reshape(B * L, -1), imgs, mask, dl, B, Hp, 16, L, 0, norm_pix, ms)
    gref = OM.forward_loss_grad(G["mae_imgs"].astype(np.float32).astype(np.float64), pred.float().cpu().numpy().astype(np.float64),
                                G["mae_mask"], norm_pix)
    assert rel(dpred.reshape(B, L, -1), torch.from_numpy(gref)) < 1e-2
    # identities: pred == target -> 0; independent of kept patches
    tgt = torch.from_numpy(OM.patchify(G["mae_imgs"])).float().cuda().bfloat16()
    l0 = V.mae_loss_fwd(tgt.reshape(B * L, -1), imgs, mask, B, Hp, 16, L, 0, False, ms)
    assert l0.item() < 1e-4


def _small_mae(norm_pix):
    from passl_b200.models.mae import MaskedAutoencoderViT
    torch.manual_seed(0)
    m = MaskedAutoencoderViT(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, decoder_embed_dim=64,
                             decoder_depth=2, decoder_num_heads=2, norm_pix_loss=norm_pix).cuda()
    for n, p in m.named_parameters():
        if p.dim() == 1 and "norm" not in n:
            torch.nn.init.normal_(p, 0, 0.1)
    return m


@pytest.mark.parametrize("norm_pix", [False, True])
def test_mae_model_fwd_bwd_vs_oracle(norm_pix):
    from oracle import vit as OV
    m = _small_mae(norm_pix)
    B, L = 6, 16
    imgs = torch.randn(B, 3, 64, 64, device="cuda")
    noise = torch.rand(B, L, device="cuda")
    for p in m.parameters():
        if p.requires_grad:
            p.grad = torch.zeros_like(p)
    loss, pred, mask = m(imgs, 0.75, noise=noise)
    loss.backward()
    torch.cuda.synchronize()
    p = OV.export_params(m)
    cfg = dict(patch=16, heads=2, dec_heads=2, depth=2, dec_depth=2, norm_pix=norm_pix)
    lr, pr, mr, ir = OV.mae_forward(imgs.cpu().double(), noise.cpu().double(), p, cfg)
    lr.backward()
    assert torch.equal(mask.cpu().double(), mr)
    assert abs(loss.item() - lr.item()) < 2e-2 * abs(lr.item()), (loss.item(), lr.item())
    assert rel(pred, pr) < 3e-2, rel(pred, pr)
    bad = []
    for name, prm in m.named_parameters():
        if not prm.requires_grad or p[name].grad is None or p[name].grad.norm() == 0:
            continue
        c = cos(prm.grad, p[name].grad)
        if c < 0.98:
            bad.append((name, round(c, 4), round(rel(prm.grad, p[name].grad), 4)))
    assert not bad, bad


def test_vit_encoder_fwd_bwd_vs_oracle():
    from oracle import vit as OV
    from passl_b200.models.vision_transformer import VisionTransformer
    import torch.nn.functional as F
    torch.manual_seed(0)
    m = VisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, qkv_bias=True, epsilon=1e-6).cuda()
    B = 5
    imgs = torch.randn(B, 3, 64, 64, device="cuda")
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
    feat = m(imgs)
    g = torch.randn_like(feat)
    feat.backward(g)
    torch.cuda.synchronize()
    p = OV.export_params(m)
    x = OV.patchify(imgs.cpu().double().bfloat16().double(), 16)
    x = F.linear(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"])
    x = torch.cat([p["cls_token"].expand(B, -1, -1), x], 1) + p["pos_embed"]
    for i in range(2):
        x = OV.block(x, p, "blocks.%d." % i, 2)
    x = F.layer_norm(x, (128,), p["norm.weight"], p["norm.bias"], 1e-6)[:, 0]
    x.backward(g.cpu().double())
    assert rel(feat, x) < 2e-2
    for name in ["cls_token", "pos_embed", "blocks.0.qkv.weight", "blocks.1.fc1.weight", "patch_embed.proj.weight", "norm.weight",
                 "blocks.0.norm1.bias", "blocks.1.proj.bias"]:
        c = cos(dict(m.named_parameters())[name].grad, p[name].grad)
        assert c > 0.98, (name, c)
