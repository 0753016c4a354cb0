"""oracle/clip.py (encode_image / encode_text / clip_forward / clip_head) against golden outputs of the reference's whole CLIP
model + CLIPHead (passl_v110/modeling/backbones/clip.py:184-338, heads/clip_head.py:27-35, executed over the paddle shim at a reduced
size, tests/golden/make_golden_models.py gen_clip_model)."""
import os

import numpy as np
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_clip_model.npz"))


def _params():
    p = {}
    for key in G.files:
        if not key.startswith("w_"):
            continue
        name, v = key[2:], torch.from_numpy(G[key])
        name = name.replace(".attn.qkv.", ".qkv.").replace(".attn.proj.", ".proj.").replace(".mlp.fc", ".fc")
        if name == "visual.patch_embed.proj.weight":          # Conv2D [W, C, p, q] -> linear on (p, q, c)
            p[name] = v.permute(0, 2, 3, 1).reshape(v.shape[0], -1)
        elif name == "visual.proj":                           # parameter [width, out_dim]
            p["visual.proj.weight"] = v.t()
        elif name == "text_projection":
            p["text.text_projection.weight"] = v.t()
        elif name == "token_embedding.weight":
            p["text.token_embedding"] = v
        elif name == "positional_embedding":
            p["text.positional_embedding"] = v
        elif name.startswith("transformer."):
            p["text." + name[len("transformer."):]] = v.t() if v.dim() == 2 else v
        elif name.startswith("ln_final."):
            p["text." + name] = v
        elif name.startswith("visual.") and v.dim() == 2 and ".blocks." in name:
            p[name] = v.t()                                   # paddle Linear [in, out] -> [out, in]
        else:
            p[name] = v
    return p


def test_clip_model_matches_reference():
    import oracle.clip as OC
    p = _params()
    cfg = dict(patch_size=int(G["cfg_vision_patch_size"]), width=int(G["cfg_vision_width"]), depth=int(G["cfg_vision_layers"]),
               num_heads=int(G["cfg_vision_width"]) // 64, pre_norm=True, text_width=int(G["cfg_transformer_width"]),
               text_layers=int(G["cfg_transformer_layers"]), text_heads=int(G["cfg_transformer_heads"]), round_pixels=False)
    img, text = torch.from_numpy(G["img"]), torch.from_numpy(G["text"])
    fi = OC.encode_image(img, p, cfg)
    ft = OC.encode_text(text, p, cfg)
    np.testing.assert_allclose(fi.numpy(), G["image_features"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ft.numpy(), G["text_features"], rtol=1e-9, atol=1e-11)
    il, tl, ls_after = OC.clip_forward(fi, ft, p["logit_scale"])
    np.testing.assert_allclose(il.numpy(), G["image_logits"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(tl.numpy(), G["text_logits"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ls_after.numpy(), G["logit_scale_after"], rtol=0, atol=0)
    o = OC.clip_head(il, tl)
    for k in ("img_loss", "text_loss", "loss"):
        np.testing.assert_allclose(o[k].item(), float(G[k]), rtol=1e-10)
