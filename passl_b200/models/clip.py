"""CLIP two-tower model on the tcgen05 kernels — mirrors passl_v110/modeling/backbones/clip.py:184-338 (`CLIP`),
backbones/vision_transformer.py:95-383 (v110 `Attention/Block/Transformer/VisionTransformer` with `pre_norm`, `proj`,
QuickGELU, `attn_mask`), heads/clip_head.py:27-35 (`CLIPHead`) and architectures/CLIPWrapper.py:26-51 (`CLIPWrapper`).

Vision tower : patch embedding (no bias) -> [class_embedding | patches] + positional_embedding -> norm_pre -> blocks (QuickGELU)
               -> norm_post(cls) -> @ proj.
Text tower   : token embedding gather + positional embedding -> causal blocks (the reference's additive -inf upper-triangular
               mask, clip.py:293-295, is the `causal` flag of the fused attention kernel) -> EOT row -> ln_final -> @ text_projection.
               (LayerNorm is per row, so pooling the EOT row before ln_final equals the reference's ln_final-then-pool.)
Loss         : image_logits = exp(logit_scale) * I_n T_n^T, text_logits = its transpose, CE(arange) both ways; logit_scale is
               read, exponentiated and clamped to [-4.6, 4.6] on the device (no D2H sync) — kernels in csrc/clip.cu.

Each tower is one autograd node; the loss is one autograd node whose backward produces dI, dT and accumulates d logit_scale.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from .. import kernels_vit as V
from ..core.param_store import compute_copy, grad_buffer
from ..nn.layers import Linear
from .vision_transformer import Block, LayerNorm, PatchEmbed, _check_unbuilt_options, _linear_bwd


def _trunc_normal_(t, std=0.02):       # vision_transformer.py(v110):31  TruncatedNormal(std=0.02)
    nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class _TowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, anchor):
        out, saved = module._run_forward(x, save=True)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.module._run_backward(ctx.saved, dout.contiguous())
        ctx.saved = None
        return None, None, None


class _Tower(nn.Module):
    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _TowerFn.apply(self, x, next(p for p in self.parameters() if p.requires_grad))
        return self._run_forward(x, save=False)[0]


class CLIPVisionTransformer(_Tower):
    """v110 backbones/vision_transformer.py:266-383 on the `proj is not None` path (what CLIP builds, clip.py:218-228)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, width=768, out_dim=512, depth=12, num_heads=12, mlp_ratio=4,
                 qkv_bias=True, pre_norm=False, proj=True, patch_bias=True, epsilon=1e-5, **kwargs):
        super().__init__()
        assert proj, "the feature-map output path (proj=None) is outside the CLIP hot path"
        _check_unbuilt_options(kwargs)
        self.width = self.num_features = width
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, width, bias=patch_bias)
        L = self.patch_embed.num_patches
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(torch.zeros(1, 1, width))
        self.positional_embedding = nn.Parameter(torch.zeros(1, L + 1, width))
        self.proj = Linear(width, out_dim, bias=False, out_fp32=True)        # reference parameter is [width, out_dim]
        self.norm_pre = LayerNorm(width, epsilon) if pre_norm else None
        self.blocks = nn.ModuleList([Block(width, num_heads, mlp_ratio, qkv_bias, epsilon, act="quick_gelu") for _ in range(depth)])
        self.norm_post = LayerNorm(width, epsilon)
        with torch.no_grad():
            self.proj.weight.normal_(0.0, scale)
            _trunc_normal_(self.positional_embedding)            # vision_transformer.py:339-340 (overrides Normal(std=scale))
            _trunc_normal_(self.class_embedding)
            for blk in self.blocks:                              # _init_weights :343-350
                for lin in (blk.qkv, blk.proj, blk.fc1, blk.fc2):
                    _trunc_normal_(lin.weight)
                    if lin.bias is not None:
                        lin.bias.zero_()

    def _run_forward(self, img, save=True):
        B = img.shape[0]
        L, D = self.patch_embed.num_patches, self.width
        pe, cols = self.patch_embed.fwd(img.float().contiguous(), save)
        x = V.token_assemble_fwd(pe, self.positional_embedding.view(L + 1, D), self.class_embedding.view(D), B, L, L + 1, D, "vit")
        cpre = None
        if self.norm_pre is not None:
            x, cpre = self.norm_pre.fwd(x)
        ctxs = []
        for blk in self.blocks:
            x, c = blk.fwd(x, B, L + 1, save)
            ctxs.append(c)
        cls = x.view(B, L + 1, D)[:, 0].contiguous()                               # x[:, 0, :]
        y, cn = self.norm_post.fwd(cls)
        feat = K.gemm(y, compute_copy(self.proj.weight), out_dtype=torch.float32)  # [B, out_dim] fp32
        return feat, (cols, cpre, ctxs, cn, y, B)

    def _run_backward(self, saved, dfeat):
        cols, cpre, ctxs, cn, y, B = saved
        L, D = self.patch_embed.num_patches, self.width
        dy = _linear_bwd(self.proj, y, K.cast_bf16(dfeat.float().contiguous()))
        dcls = self.norm_post.bwd(cn, dy)
        d = torch.zeros((B, L + 1, D), dtype=torch.bfloat16, device=dfeat.device)
        d[:, 0].copy_(dcls)
        d = d.view(B * (L + 1), D)
        for blk, c in zip(reversed(self.blocks), reversed(ctxs)):
            d = blk.bwd(c, d)
        if self.norm_pre is not None:
            d = self.norm_pre.bwd(cpre, d)
        dpe = V.token_assemble_bwd(d, B, L, L + 1, D, "vit", acc_tok=grad_buffer(self.class_embedding).view(D),
                                   acc_pos=grad_buffer(self.positional_embedding).view(L + 1, D))
        self.patch_embed.bwd(cols, dpe)


class CLIPTextTransformer(_Tower):
    """Text side of clip.py:230-253,299-314: token_embedding, positional_embedding, causal Transformer, ln_final, text_projection."""

    def __init__(self, embed_dim=512, context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8,
                 transformer_layers=12, qkv_bias=True, epsilon=1e-5):
        super().__init__()
        W = transformer_width
        self.context_length, self.vocab_size, self.width, self.depth = context_length, vocab_size, W, transformer_layers
        self.token_embedding = nn.Parameter(torch.empty(vocab_size, W))
        self.positional_embedding = nn.Parameter(torch.empty(context_length, W))
        self.blocks = nn.ModuleList([Block(W, transformer_heads, 4, qkv_bias, epsilon, act="quick_gelu", causal=True)
                                     for _ in range(transformer_layers)])
        self.ln_final = LayerNorm(W, epsilon=1e-5)
        self.text_projection = Linear(W, embed_dim, bias=False, out_fp32=True)   # reference parameter is [width, embed_dim]
        # clip.py:259-288
        proj_std = (W ** -0.5) * (2 * transformer_layers)      # sic: the reference multiplies (clip.py:279-280)
        attn_std, fc_std = W ** -0.5, (2 * W) ** -0.5
        with torch.no_grad():
            self.token_embedding.normal_(0.0, 0.02)
            self.positional_embedding.normal_(0.0, 0.01)
            self.text_projection.weight.normal_(0.0, W ** -0.5)
            for blk in self.blocks:
                blk.proj.weight.normal_(0.0, proj_std)
                blk.qkv.weight.normal_(0.0, attn_std)
                blk.fc1.weight.normal_(0.0, fc_std)
                blk.fc2.weight.normal_(0.0, proj_std)

    def _run_forward(self, text, save=True):
        B, L = text.shape
        assert L == self.context_length and text.dtype == torch.int64
        text = text.contiguous()
        x = V.embedding_fwd(text, self.token_embedding, self.positional_embedding)
        ctxs = []
        for blk in self.blocks:
            x, c = blk.fwd(x, B, L, save)
            ctxs.append(c)
        eot, idx = V.eot_gather_fwd(text, x)
        y, cn = self.ln_final.fwd(eot)
        feat = K.gemm(y, compute_copy(self.text_projection.weight), out_dtype=torch.float32)
        return feat, (text, ctxs, idx, cn, y)

    def _run_backward(self, saved, dfeat):
        text, ctxs, idx, cn, y = saved
        dy = _linear_bwd(self.text_projection, y, K.cast_bf16(dfeat.float().contiguous()))
        deot = self.ln_final.bwd(cn, dy)
        d = V.eot_gather_bwd(idx, deot, self.context_length)
        for blk, c in zip(reversed(self.blocks), reversed(ctxs)):
            d = blk.bwd(c, d)
        V.embedding_bwd(text, d, dtable=grad_buffer(self.token_embedding), dpos=grad_buffer(self.positional_embedding))


def _pad_rows(x, n_pad):
    if x.shape[0] == n_pad:
        return x
    out = torch.zeros((n_pad, x.shape[1]), dtype=x.dtype, device=x.device)
    out[:x.shape[0]].copy_(x)
    return out


class _ClipLossFn(torch.autograd.Function):
    """normalise -> C = I_n T_n^T (tcgen05 GEMM, fp32) -> symmetric CE with exp(logit_scale) applied on the device.
    Batches that are not a multiple of 8 are zero-padded for the GEMM; the CE kernels only see the valid n x n block."""

    @staticmethod
    def forward(ctx, img_f, txt_f, logit_scale, clamp):
        n = img_f.shape[0]
        n_pad = (n + 7) // 8 * 8
        _, ib, iinv = K.l2norm_fwd(img_f.contiguous(), mode="clip", want_bf16=True)
        _, tb, tinv = K.l2norm_fwd(txt_f.contiguous(), mode="clip", want_bf16=True)
        ibp, tbp = _pad_rows(ib, n_pad), _pad_rows(tb, n_pad)
        C = K.gemm(ibp, tbp, out_dtype=torch.float32)                     # [n_pad, n_pad] cosine similarities
        out3, ws = V.clip_ce_fwd(C, logit_scale.data, n=n, clamp=clamp)
        ctx.saved = (ib, tb, ibp, tbp, iinv, tinv, C, ws, logit_scale, n)
        return out3[0], out3[1], out3[2]

    @staticmethod
    def backward(ctx, d_img, d_txt, d_loss):
        # img_loss / text_loss are reported values (clip_head.py:33-34); the training signal is d(loss)
        ib, tb, ibp, tbp, iinv, tinv, C, ws, logit_scale, n = ctx.saved
        ctx.saved = None
        dls = grad_buffer(logit_scale) if logit_scale.requires_grad else None
        dC = V.clip_ce_bwd(C, ws, n=n, dloss=d_loss.contiguous().float().reshape(1), dlogit_scale=dls)
        dI_n = K.gemm(dC, tbp, b_t=True, out_dtype=torch.float32)[:n]               # dC   @ T_n
        dT_n = K.gemm(dC, ibp, a_t=True, b_t=True, out_dtype=torch.float32)[:n]     # dC^T @ I_n
        yi = torch.empty(ib.shape, dtype=torch.float32, device=ib.device)
        yt = torch.empty(tb.shape, dtype=torch.float32, device=tb.device)
        K.cast_f32(ib, yi)
        K.cast_f32(tb, yt)
        dI, _ = K.l2norm_bwd(dI_n, yi, iinv, mode="clip")
        dT, _ = K.l2norm_bwd(dT_n, yt, tinv, mode="clip")
        return dI, dT, None, None


class CLIPHead(nn.Module):
    """clip_head.py:22-35.  `forward(img_logits, text_logits, img_labels, text_labels)` keeps the reference signature for
    materialised logits (debug path); the hot path calls `forward_fused(image_features, text_features, logit_scale)`."""

    def forward_fused(self, image_features, text_features, logit_scale, clamp=True):
        img_loss, text_loss, loss = _ClipLossFn.apply(image_features, text_features, logit_scale, clamp)
        return {"img_loss": img_loss, "text_loss": text_loss, "loss": loss}

    def forward(self, img_logits, text_logits, img_labels, text_labels):
        from ..loss.contrastive import logits_cross_entropy
        img_loss = logits_cross_entropy(img_logits, img_labels)
        text_loss = logits_cross_entropy(text_logits, text_labels)
        return {"img_loss": img_loss, "text_loss": text_loss, "loss": img_loss + text_loss}


class CLIP(nn.Module):
    """clip.py:184-338 with the ViT vision tower (the `vision_layers: int` branch; ModifiedResNet is out of scope)."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, pre_norm, proj, patch_bias,
                 context_length, vocab_size, transformer_width, transformer_heads, transformer_layers, qkv_bias):
        super().__init__()
        assert not isinstance(vision_layers, (tuple, list)), "ModifiedResNet vision tower is not part of the hot path"
        self.context_length = context_length
        self.visual = CLIPVisionTransformer(img_size=image_resolution, patch_size=vision_patch_size, width=vision_width,
                                            out_dim=embed_dim, depth=vision_layers, num_heads=vision_width // 64,
                                            pre_norm=pre_norm, proj=proj, patch_bias=patch_bias)
        self.text = CLIPTextTransformer(embed_dim, context_length, vocab_size, transformer_width, transformer_heads,
                                        transformer_layers, qkv_bias)
        self.logit_scale = nn.Parameter(torch.full((1,), float(np.log(1 / 0.07))))

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, text):
        return self.text(text)

    def forward(self, image, text, is_train=True):
        """Returns (image_features, text_features) — the un-normalised tower outputs; `CLIPHead.forward_fused` consumes them
        (the reference returns the two [n, n] logit matrices here, clip.py:320-338; they are never materialised twice)."""
        return self.encode_image(image), self.encode_text(text)


class CLIPWrapper(nn.Module):
    """architectures/CLIPWrapper.py:26-65: `forward(image, text, mode='train') -> {'img_loss','text_loss','loss'}`."""

    def __init__(self, architecture=None, head=None):
        super().__init__()
        from ..modeling.registry import build_backbone, build_head
        self.model = architecture if isinstance(architecture, nn.Module) else build_backbone(dict(architecture))
        self.head = head if isinstance(head, nn.Module) else build_head(dict(head or {"name": "CLIPHead"}))

    def train_iter(self, *inputs, **kwargs):
        image, text = inputs
        img_f, txt_f = self.model(image, text, is_train=True)
        return self.head.forward_fused(img_f, txt_f, self.model.logit_scale)

    def forward(self, *inputs, mode='train', **kwargs):
        if mode == 'train':
            return self.train_iter(*inputs, **kwargs)
        elif mode == 'extract':
            return self.model(*inputs)
        raise Exception("No such mode: {}".format(mode))
