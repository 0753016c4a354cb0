"""Vision Transformer blocks on the tcgen05 kernels — mirrors passl/models/vision_transformer.py:84-363
(`Mlp`, `Attention`, `Block`, `PatchEmbed`, `VisionTransformer`, `ViT_base_patch16_224`).

Tokens live as a bf16 [B*N, D] residual stream.  One Block = 2 LayerNorm kernels, 4 GEMMs with fused epilogues (qkv bias;
proj bias + residual; fc1 bias + exact GELU + pre-activation copy; fc2 bias + residual) and the fused attention kernel.
Backward is hand-written: dgrad GEMMs (GELU' fused in the fc2-dgrad epilogue), wgrad GEMMs accumulated into the flat
gradient buffer, attention backward with recomputed probabilities, LayerNorm backward with the residual-gradient add fused.
Dropout / drop-path rates are 0 in every pretrain config of the hot path and are not implemented.
"""
import torch
import torch.nn as nn

from .. import kernels as K
from .. import kernels_vit as V
from ..core.param_store import compute_copy, grad_buffer
from ..nn.layers import Linear


class LayerNorm(nn.Module):
    def __init__(self, dim, epsilon=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.eps = epsilon

    def fwd(self, x):
        y, mean, rstd = V.layernorm_fwd(x, self.weight, self.bias, self.eps)
        return y, (x, mean, rstd)

    def bwd(self, ctx, dy, dres=None):
        x, mean, rstd = ctx
        tr = self.weight.requires_grad
        dx, _ = V.layernorm_bwd(x, dy, self.weight, mean, rstd, dgamma=grad_buffer(self.weight) if tr else None,
                                dbeta=grad_buffer(self.bias) if tr else None, dres=dres)
        return dx


def _linear_bwd(lin, x, dy, need_dx=True, aux=None, aux_mode="gelu_grad"):
    """Shared dgrad / wgrad / bias-grad of y = x W^T + b for bf16 [T, *] operands."""
    if lin.weight.requires_grad:
        K.gemm(dy, x, a_t=True, b_t=True, out=grad_buffer(lin.weight), accumulate=True,
               splits=K.wgrad_splits(lin.cout, lin.cin, x.shape[0]))
        if lin.bias is not None:
            K.colsum_accumulate(dy, grad_buffer(lin.bias))
    if not need_dx:
        return None
    return K.gemm(dy, compute_copy(lin.weight), b_t=True, aux=aux, aux_mode_name=aux_mode)


class Block(nn.Module):
    """Pre-LN transformer block (vision_transformer.py:159-206)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, epsilon=1e-6, act="gelu", causal=False):
        super().__init__()
        assert dim % num_heads == 0
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.causal = causal
        self.act = act
        hidden = int(dim * mlp_ratio)
        self.norm1 = LayerNorm(dim, epsilon)
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)
        self.norm2 = LayerNorm(dim, epsilon)
        self.fc1 = Linear(dim, hidden)
        self.fc2 = Linear(hidden, dim)
        for lin in (self.qkv, self.proj, self.fc1, self.fc2):          # timm / MAE init: xavier_uniform, zero bias
            nn.init.xavier_uniform_(lin.weight)
            if lin.bias is not None:
                nn.init.zeros_(lin.bias)

    def fwd(self, x, B, N, save=True):
        h1, c1 = self.norm1.fwd(x)
        qkv = K.gemm(h1, compute_copy(self.qkv.weight), bias=self.qkv.bias)
        a, lse = V.attention_fwd(qkv, B, N, self.num_heads, self.head_dim, scale=self.scale, causal=self.causal)
        x1 = K.gemm(a, compute_copy(self.proj.weight), bias=self.proj.bias, residual=x)
        h2, c2 = self.norm2.fwd(x1)
        u = torch.empty((x.shape[0], self.fc1.cout), dtype=torch.bfloat16, device=x.device) if save else None
        hf = K.gemm(h2, compute_copy(self.fc1.weight), bias=self.fc1.bias, act=self.act, preact_out=u)
        x2 = K.gemm(hf, compute_copy(self.fc2.weight), bias=self.fc2.bias, residual=x1)
        ctx = (c1, h1, qkv, a, lse, c2, h2, u, hf, B, N) if save else None
        return x2, ctx

    def bwd(self, ctx, dx2):
        c1, h1, qkv, a, lse, c2, h2, u, hf, B, N = ctx
        # MLP branch: x2 = x1 + fc2(act(fc1(LN2(x1))))
        du = _linear_bwd(self.fc2, hf, dx2, aux=u, aux_mode="gelu_grad" if self.act == "gelu" else "quick_gelu_grad")
        dh2 = _linear_bwd(self.fc1, h2, du)
        dx1 = self.norm2.bwd(c2, dh2, dres=dx2)                      # + gradient through the residual connection
        # attention branch: x1 = x + proj(attn(qkv(LN1(x))))
        da = _linear_bwd(self.proj, a, dx1)
        dqkv = V.attention_bwd(qkv, da, a, lse, B, N, self.num_heads, self.head_dim, scale=self.scale, causal=self.causal)
        dh1 = _linear_bwd(self.qkv, h1, dqkv)
        return self.norm1.bwd(c1, dh1, dres=dx1)


class PatchEmbed(nn.Module):
    """Conv2D(k = s = patch) as im2col + GEMM (vision_transformer.py:209-249); weight [embed, p, p, c] flattened (p,q,c)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.kdim = patch_size * patch_size * in_chans
        self.proj = Linear(self.kdim, embed_dim, bias=bias)
        # mae.py:133-137: xavier_uniform on the weight viewed as [embed, -1]
        nn.init.xavier_uniform_(self.proj.weight)

    def fwd(self, img, save=True):
        B, C, H, W = img.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        p = self.patch_size[0]
        cols, _, _ = K.im2col_nchw(img.contiguous(), p, p, p, 0, self.kdim)
        out = K.gemm(cols, compute_copy(self.proj.weight), bias=self.proj.bias)
        return out, (cols if save else None)

    def bwd(self, cols, dout):
        _linear_bwd(self.proj, cols, dout, need_dx=False)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, img, anchor):
        out, saved = module._run_forward(img, save=True)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.module._run_backward(ctx.saved, dout.contiguous())
        ctx.saved = None
        return None, None, None


def _check_unbuilt_options(kwargs):
    """Constructor keys of the reference ViTs (vision_transformer.py:256-275) whose only built value is the pre-training default:
    dropout / stochastic depth are 0 on every self-supervised recipe here, so a non-zero request must fail instead of being
    dropped on the floor.  `norm_layer` is accepted (always LayerNorm with the `epsilon` given)."""
    inert = {"drop_rate": (0, 0.0), "attn_drop_rate": (0, 0.0), "drop_path_rate": (0, 0.0), "qk_scale": (None,),
             "representation_size": (None,)}
    for k, v in kwargs.items():
        if k == "norm_layer":
            continue
        if k not in inert:
            raise TypeError("unexpected keyword argument %r" % k)
        if v not in inert[k]:
            raise NotImplementedError("%s=%r is not built (only %r)" % (k, v, inert[k][0]))


class VisionTransformer(nn.Module):
    """passl/models/vision_transformer.py:252-363 (feature extractor: returns the cls token after the final norm)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, class_num=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4, qkv_bias=False, epsilon=1e-5, learnable_pos=True, **kwargs):
        super().__init__()
        assert class_num <= 0, "the classification head is outside the self-supervised hot path"
        _check_unbuilt_options(kwargs)
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        L = self.patch_embed.num_patches
        self.pos_embed = nn.Parameter(torch.zeros(1, L + 1, embed_dim), requires_grad=learnable_pos)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias, epsilon) for _ in range(depth)])
        self.norm = LayerNorm(embed_dim, epsilon)
        nn.init.normal_(self.pos_embed, std=0.02)

    def _run_forward(self, img, save=True):
        B = img.shape[0]
        L, D = self.patch_embed.num_patches, self.embed_dim
        pe, cols = self.patch_embed.fwd(img, save)
        x = V.token_assemble_fwd(pe, self.pos_embed.view(L + 1, D), self.cls_token.view(D), B, L, L + 1, D, "vit")
        ctxs = []
        for blk in self.blocks:
            x, c = blk.fwd(x, B, L + 1, save)
            ctxs.append(c)
        y, cn = self.norm.fwd(x)
        feat = y.view(B, L + 1, D)[:, 0].contiguous()               # x[:, 0]
        return feat, (cols, ctxs, cn, B)

    def _run_backward(self, saved, dfeat):
        cols, ctxs, cn, B = saved
        L, D = self.patch_embed.num_patches, self.embed_dim
        dy = torch.zeros((B, L + 1, D), dtype=torch.bfloat16, device=dfeat.device)
        dy[:, 0].copy_(dfeat)
        d = self.norm.bwd(cn, dy.view(B * (L + 1), D))
        for blk, c in zip(reversed(self.blocks), reversed(ctxs)):
            d = blk.bwd(c, d)
        dpe = V.token_assemble_bwd(d, B, L, L + 1, D, "vit", acc_tok=grad_buffer(self.cls_token).view(D),
                                   acc_pos=grad_buffer(self.pos_embed).view(L + 1, D) if self.pos_embed.requires_grad else None)
        self.patch_embed.bwd(cols, dpe)

    def forward(self, img):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _EncoderFn.apply(self, img, self.cls_token)
        out, _ = self._run_forward(img, save=False)
        return out


def ViT_base_patch16_224(**kwargs):
    """vision_transformer.py:432-443"""
    kw = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True, epsilon=1e-6)
    kw.update(kwargs)
    return VisionTransformer(**kw)
