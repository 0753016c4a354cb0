"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the CLIP path: torch-CPU float64 restatement of
passl_v110/modeling/backbones/clip.py:184-338 (CLIP.encode_image / encode_text / forward / clip_logit_scale),
backbones/vision_transformer.py:95-183 (v110 Attention with additive attn_mask, Block with QuickGELU), :266-371
(VisionTransformer.forward_features on the `proj` path), base_transformer.py:25-28 (QuickGELU), heads/clip_head.py:27-35
(CLIPHead) and architectures/CLIPWrapper.py:45-51 (labels = arange).  Driven by a parameter dict exported from the CUDA module
(oracle.vit.export_params: 2-D weights rounded to bf16 = what the tensor cores multiply; weights stored [out, in]).
Pinned against the reference's own source run over the paddle shim: CLIPHead (reference_heads.npz), the v110 Block with / without
the causal mask (reference_clip_block.npz) and the whole CLIP model + head at a reduced size (reference_clip_model.npz: both towers,
logits, losses, logit_scale clamp) — tests/test_oracle_vit_cpu.py, tests/test_oracle_clip_model_cpu.py."""
import math

import torch
import torch.nn.functional as F


def quick_gelu(x):
    """base_transformer.py:25-28"""
    return x * torch.sigmoid(1.702 * x)


def block(x, p, pre, num_heads, eps=1e-5, attn_mask=None):
    """vision_transformer.py(v110):141-183: x + attn(norm1(x)); x + mlp(norm2(x)), act = QuickGELU."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps)
    qkv = F.linear(h, p[pre + "qkv.weight"], p.get(pre + "qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-1, -2)) * (C // num_heads) ** -0.5
    if attn_mask is not None:
        attn = attn + attn_mask
    attn = torch.softmax(attn, dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(a, p[pre + "proj.weight"], p[pre + "proj.bias"])
    h = F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
    h = quick_gelu(F.linear(h, p[pre + "fc1.weight"], p[pre + "fc1.bias"]))
    return x + F.linear(h, p[pre + "fc2.weight"], p[pre + "fc2.bias"])


def build_attention_mask(length, dtype=torch.float64):
    """clip.py:293-295: upper triangle (diagonal excluded) = -inf"""
    return torch.triu(torch.full((length, length), -math.inf, dtype=dtype), 1)


def encode_image(img, p, cfg, pre="visual."):
    """vision_transformer.py(v110):352-365 with proj (patch conv as unfold + matmul, weight [W, p, p, c] flattened (p, q, c))."""
    B = img.shape[0]
    ps, W = cfg["patch_size"], cfg["width"]
    g = img.shape[2] // ps
    x = img.reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 3, 5, 1).reshape(B, g * g, ps * ps * 3)     # (p, q, c) order
    if cfg.get("round_pixels", True):
        x = x.bfloat16().to(img.dtype)                                                                # im2col emits bf16 operands
    x = F.linear(x, p[pre + "patch_embed.proj.weight"], p.get(pre + "patch_embed.proj.bias"))
    x = torch.cat([p[pre + "class_embedding"].expand(B, -1, -1), x], dim=1) + p[pre + "positional_embedding"]
    if cfg.get("pre_norm", False):
        x = F.layer_norm(x, (W,), p[pre + "norm_pre.weight"], p[pre + "norm_pre.bias"], 1e-5)
    for i in range(cfg["depth"]):
        x = block(x, p, pre + "blocks.%d." % i, cfg["num_heads"])
    x = F.layer_norm(x[:, 0, :], (W,), p[pre + "norm_post.weight"], p[pre + "norm_post.bias"], 1e-5)
    return F.linear(x, p[pre + "proj.weight"])


def encode_text(text, p, cfg, pre="text."):
    """clip.py:299-314"""
    W, L = cfg["text_width"], text.shape[1]
    x = p[pre + "token_embedding"][text] + p[pre + "positional_embedding"]
    mask = build_attention_mask(L, x.dtype)
    for i in range(cfg["text_layers"]):
        x = block(x, p, pre + "blocks.%d." % i, cfg["text_heads"], attn_mask=mask)
    x = F.layer_norm(x, (W,), p[pre + "ln_final.weight"], p[pre + "ln_final.bias"], 1e-5)
    idx = text.argmax(dim=-1)
    x = x[torch.arange(x.shape[0]), idx]
    return F.linear(x, p[pre + "text_projection.weight"])


def clip_forward(image_features, text_features, logit_scale):
    """clip.py:320-338 -> (image_logits, text_logits, clamped logit_scale)"""
    i = image_features / image_features.norm(dim=-1, keepdim=True)
    t = text_features / text_features.norm(dim=-1, keepdim=True)
    s = logit_scale.exp()
    return (s * i) @ t.t(), (s * t) @ i.t(), logit_scale.detach().clamp(-4.6, 4.6)


def clip_head(img_logits, text_logits):
    """clip_head.py:27-35 with CLIPWrapper.py:47-48 labels"""
    n = img_logits.shape[0]
    labels = torch.arange(n)
    img_loss = F.cross_entropy(img_logits, labels)
    text_loss = F.cross_entropy(text_logits, labels)
    return {"img_loss": img_loss, "text_loss": text_loss, "loss": img_loss + text_loss}


def clip_train_iter(image, text, p, cfg):
    fi = encode_image(image, p, cfg)
    ft = encode_text(text, p, cfg)
    il, tl, ls = clip_forward(fi, ft, p["logit_scale"])
    out = clip_head(il, tl)
    out.update(image_features=fi, text_features=ft, logit_scale_after=ls)
    return out
