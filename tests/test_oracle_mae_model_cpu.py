"""oracle/vit.py `mae_forward` against golden outputs of the reference's whole MaskedAutoencoderViT
(passl/models/mae.py:37-290 executed over the paddle shim at a reduced size, tests/golden/make_golden_models.py gen_mae_model):
loss, prediction, mask — with and without norm_pix_loss."""
import os

import numpy as np
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_mae_model.npz"))


def _params():
    p = {}
    for key in G.files:
        if not key.startswith("w_"):
            continue
        name, v = key[2:], torch.from_numpy(G[key])
        name = name.replace(".attn.qkv.", ".qkv.").replace(".attn.proj.", ".proj.").replace(".mlp.fc", ".fc")
        if name == "patch_embed.proj.weight":                 # Conv2D [E, C, p, q] -> linear on (p, q, c)
            v = v.permute(0, 2, 3, 1).reshape(v.shape[0], -1)
        elif v.dim() == 2:                                    # paddle Linear [in, out] -> [out, in]
            v = v.t()
        p[name] = v
    return p


def test_mae_forward_matches_reference_model():
    import oracle.vit as OV
    p = _params()
    imgs, noise = torch.from_numpy(G["imgs"]), torch.from_numpy(G["noise"])
    for npl in (0, 1):
        cfg = dict(patch=8, heads=2, dec_heads=2, depth=2, dec_depth=1, norm_pix=bool(npl), round_pixels=False)
        loss, pred, mask, ids_restore = OV.mae_forward(imgs, noise, p, cfg, mask_ratio=0.75)
        np.testing.assert_allclose(loss.item(), float(G["loss%d" % npl]), rtol=1e-10)
        np.testing.assert_allclose(pred.numpy(), G["pred%d" % npl], rtol=1e-8, atol=1e-10)
        assert np.array_equal(mask.numpy(), G["mask%d" % npl])                       # 0 keep / 1 remove, bit-exact
        assert mask.sum().item() == 4 * (16 - int(16 * 0.25))
