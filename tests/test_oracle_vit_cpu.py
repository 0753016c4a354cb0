"""Cross-oracle for the transformer block: oracle/vit.py `block` (restating passl/models/vision_transformer.py:116-206) and
oracle/clip.py `block` (QuickGELU, additive causal mask) against torch.nn.TransformerEncoderLayer(norm_first=True) — an independent
implementation of the same pre-LN block — with copied weights in float64."""
import torch
import torch.nn as nn


def _layer_and_params(D, H, act, pre="b."):
    torch.manual_seed(0)
    layer = nn.TransformerEncoderLayer(D, H, dim_feedforward=4 * D, dropout=0.0, activation=act, batch_first=True,
                                       norm_first=True, layer_norm_eps=1e-6).double().eval()
    for prm in layer.parameters():
        nn.init.normal_(prm, std=0.2)
    p = {pre + "norm1.weight": layer.norm1.weight, pre + "norm1.bias": layer.norm1.bias,
         pre + "qkv.weight": layer.self_attn.in_proj_weight, pre + "qkv.bias": layer.self_attn.in_proj_bias,
         pre + "proj.weight": layer.self_attn.out_proj.weight, pre + "proj.bias": layer.self_attn.out_proj.bias,
         pre + "norm2.weight": layer.norm2.weight, pre + "norm2.bias": layer.norm2.bias,
         pre + "fc1.weight": layer.linear1.weight, pre + "fc1.bias": layer.linear1.bias,
         pre + "fc2.weight": layer.linear2.weight, pre + "fc2.bias": layer.linear2.bias}
    return layer, {k: v.detach() for k, v in p.items()}


def test_vit_block_matches_torch_encoder_layer():
    import oracle.vit as OV
    layer, p = _layer_and_params(64, 4, "gelu")
    x = torch.randn(3, 17, 64, dtype=torch.float64)
    with torch.no_grad():
        ref = layer(x)
    got = OV.block(x, p, "b.", 4, eps=1e-6)
    torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-9)


def test_clip_block_quickgelu_causal_matches_torch_encoder_layer():
    import oracle.clip as OC
    layer, p = _layer_and_params(64, 2, OC.quick_gelu)
    x = torch.randn(2, 11, 64, dtype=torch.float64)
    mask = OC.build_attention_mask(11)
    with torch.no_grad():
        ref = layer(x, src_mask=mask)
    got = OC.block(x, p, "b.", 2, eps=1e-6, attn_mask=mask)
    torch.testing.assert_close(got, ref, rtol=1e-9, atol=1e-9)


def test_vit_block_matches_reference_block_class():
    """oracle/vit.py `block` against golden vectors from the reference's own `Block` / `Attention` / `Mlp` classes
    (passl/models/vision_transformer.py:84-206 run over the paddle shim, tests/golden/make_golden_models.py)."""
    import os
    import numpy as np
    import oracle.vit as OV
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vit_block.npz"))
    names = {"norm1": "norm1", "norm2": "norm2", "qkv": "attn.qkv", "proj": "attn.proj", "fc1": "mlp.fc1", "fc2": "mlp.fc2"}
    for tag in ("a", "b"):
        p = {}
        for mine, ref in names.items():
            w = torch.from_numpy(g["%s_%s.weight" % (tag, ref)])
            p["b.%s.weight" % mine] = w.t() if w.dim() == 2 else w          # paddle Linear weight is [in, out]
            p["b.%s.bias" % mine] = torch.from_numpy(g["%s_%s.bias" % (tag, ref)])
        y = OV.block(torch.from_numpy(g[tag + "_x"]), p, "b.", int(g[tag + "_heads"]), eps=1e-6)
        np.testing.assert_allclose(y.numpy(), g[tag + "_y"], rtol=1e-10, atol=1e-12)


def test_clip_block_matches_reference_v110_block_class():
    """oracle/clip.py `block` (QuickGELU, additive mask) against golden vectors from the reference's v110 `Block`
    (passl_v110/modeling/backbones/vision_transformer.py:141-183) with the causal mask of clip.py:293-295 and without a mask."""
    import os
    import numpy as np
    import oracle.clip as OC
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_clip_block.npz"))
    names = {"norm1": "norm1", "norm2": "norm2", "qkv": "attn.qkv", "proj": "attn.proj", "fc1": "mlp.fc1", "fc2": "mlp.fc2"}
    for tag in ("causal", "plain"):
        p = {}
        for mine, ref in names.items():
            w = torch.from_numpy(g["%s_%s.weight" % (tag, ref)])
            p["b.%s.weight" % mine] = w.t() if w.dim() == 2 else w
            p["b.%s.bias" % mine] = torch.from_numpy(g["%s_%s.bias" % (tag, ref)])
        x = torch.from_numpy(g[tag + "_x"])
        mask = OC.build_attention_mask(x.shape[1]) if tag == "causal" else None
        y = OC.block(x, p, "b.", int(g[tag + "_heads"]), eps=1e-5, attn_mask=mask)
        np.testing.assert_allclose(y.numpy(), g[tag + "_y"], rtol=1e-10, atol=1e-12)
