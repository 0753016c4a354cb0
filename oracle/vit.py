"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the ViT block and the MAE model: torch-CPU float64 restatement of
passl/models/vision_transformer.py:84-206 (Mlp :84-113, Attention :116-156, Block :159-206) and
passl/models/mae.py:214-290 (forward_encoder, forward_decoder, forward_loss, forward) driven by a parameter dict exported from
the CUDA module (weights rounded to bf16 = what the tensor cores multiply).  Paddle semantics restated: nn.Linear y = xW+b
(weights here are stored [out, in]), nn.LayerNorm(epsilon), nn.GELU exact erf, softmax over the last axis, Tensor.var unbiased.
The masking noise is an input (Paddle's RNG stream is not reproducible).  patchify / random_masking / forward_loss are pinned
against the reference source through oracle/mae.py + tests/golden (numpy); this file is the differentiable torch twin, itself
pinned against the reference's `Block` class (reference_vit_block.npz) and whole `MaskedAutoencoderViT` (reference_mae_model.npz)
run over the paddle shim (tests/test_oracle_vit_cpu.py, tests/test_oracle_mae_model_cpu.py)."""
import torch
import torch.nn.functional as F


def export_params(module, dtype=torch.float64):
    out = {}
    for name, t in module.named_parameters():
        v = t.detach().float().cpu()
        if v.dim() == 2:
            v = v.bfloat16().float()
        out[name] = v.to(dtype).requires_grad_(t.requires_grad)
    return out


def block(x, p, pre, num_heads, eps=1e-6):
    """vision_transformer.py:203-206: x + attn(norm1(x)); x + mlp(norm2(x))."""
    B, N, C = x.shape
    h = F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps)
    qkv = F.linear(h, p[pre + "qkv.weight"], p.get(pre + "qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-1, -2)) * (C // num_heads) ** -0.5
    attn = torch.softmax(attn, dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(a, p[pre + "proj.weight"], p[pre + "proj.bias"])
    h = F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
    h = F.gelu(F.linear(h, p[pre + "fc1.weight"], p[pre + "fc1.bias"]))
    return x + F.linear(h, p[pre + "fc2.weight"], p[pre + "fc2.bias"])


def patchify(imgs, p):
    n = imgs.shape[0]
    h = w = imgs.shape[2] // p
    x = imgs.reshape(n, 3, h, p, w, p)
    return torch.einsum('nchpwq->nhwpqc', x).reshape(n, h * w, p * p * 3)


def mae_forward(imgs, noise, p, cfg, mask_ratio=0.75):
    """cfg: dict(patch, heads, dec_heads, depth, dec_depth, norm_pix).  Returns (loss, pred, mask, ids_restore)."""
    P = cfg["patch"]
    x = patchify(imgs.to(torch.float64).bfloat16().double() if cfg.get("round_pixels", True) else imgs, P)   # conv k=s=16 == linear on (p,q,c)
    x = F.linear(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"])
    x = x + p["pos_embed"][:, 1:, :]
    N, L, D = x.shape
    len_keep = int(L * (1 - mask_ratio))
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :len_keep]
    x = torch.gather(x, 1, ids_keep.unsqueeze(-1).expand(-1, -1, D))
    mask = torch.ones(N, L, dtype=x.dtype)
    mask[:, :len_keep] = 0
    mask = torch.gather(mask, 1, ids_restore)
    cls = p["cls_token"] + p["pos_embed"][:, :1, :]
    x = torch.cat([cls.expand(N, -1, -1), x], 1)
    for i in range(cfg["depth"]):
        x = block(x, p, "blocks.%d." % i, cfg["heads"])
    x = F.layer_norm(x, (D,), p["norm.weight"], p["norm.bias"], 1e-6)
    x = F.linear(x, p["decoder_embed.weight"], p["decoder_embed.bias"])
    Dd = x.shape[-1]
    mask_tokens = p["mask_token"].expand(N, L + 1 - x.shape[1], -1)
    x_ = torch.cat([x[:, 1:, :], mask_tokens], 1)
    x_ = torch.gather(x_, 1, ids_restore.unsqueeze(-1).expand(-1, -1, Dd))
    x = torch.cat([x[:, :1, :], x_], 1) + p["decoder_pos_embed"]
    for i in range(cfg["dec_depth"]):
        x = block(x, p, "decoder_blocks.%d." % i, cfg["dec_heads"])
    x = F.layer_norm(x, (Dd,), p["decoder_norm.weight"], p["decoder_norm.bias"], 1e-6)
    pred = F.linear(x, p["decoder_pred.weight"], p["decoder_pred.bias"])[:, 1:, :]
    target = patchify(imgs.to(torch.float64), P)
    if cfg["norm_pix"]:
        mean = target.mean(-1, keepdim=True)
        var = target.var(-1, keepdim=True)                     # unbiased
        target = (target - mean) / (var + 1.e-6) ** .5
    loss = ((pred - target) ** 2).mean(-1)
    loss = (loss * mask).sum() / mask.sum()
    return loss, pred, mask, ids_restore
