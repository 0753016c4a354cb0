"""Masked Autoencoder ViT on the tcgen05 kernels — mirrors passl/models/mae.py:37-290, :331-344 (`MaskedAutoencoderViT`,
`mae_vit_base_patch16` = encoder 768 x 12 x 12 heads, decoder 512 x 8 x 16 heads, patch 16).

`forward(imgs, mask_ratio=0.75, noise=None) -> (loss, pred, mask)` like the reference; `noise` ([B, L] uniform) may be supplied
because Paddle's RNG stream is not reproducible — otherwise it is drawn on the device.  The whole model is one
autograd node; `loss.backward()` runs the hand-written backward and accumulates into the flat gradient buffer.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from .. import kernels_vit as V
from ..core.param_store import compute_copy, grad_buffer
from ..nn.layers import Linear
from .vision_transformer import Block, LayerNorm, PatchEmbed, _linear_bwd


def get_2d_sincos_pos_embed(embed_dim, grid_size, cls_token=False):
    """passl/models/utils/pos_embed.py:31-82 (numpy, float32 omega / grid; golden-checked bit-exact in tests)."""
    def _1d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float32)
        omega /= dim / 2.
        omega = 1. / 10000 ** omega
        out = np.einsum('m,d->md', pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    grid_h = np.arange(grid_size, dtype=np.float32)
    grid_w = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, grid_size, grid_size])
    emb = np.concatenate([_1d(embed_dim // 2, grid[0]), _1d(embed_dim // 2, grid[1])], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return emb


class _MAEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, imgs, noise, mask_ratio, anchor):
        loss, pred, mask, saved = module._run_forward(imgs, noise, mask_ratio, save=True)
        ctx.module, ctx.saved = module, saved
        ctx.mark_non_differentiable(pred, mask)
        return loss, pred, mask

    @staticmethod
    def backward(ctx, dloss, _dp, _dm):
        ctx.module._run_backward(ctx.saved, dloss.contiguous().float())
        ctx.saved = None
        return None, None, None, None, None


class MaskedAutoencoderViT(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4., norm_pix_loss=False, epsilon=1e-6):
        super().__init__()
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        L = self.patch_embed.num_patches
        self.embed_dim, self.decoder_embed_dim, self.in_chans = embed_dim, decoder_embed_dim, in_chans
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, L + 1, embed_dim), requires_grad=False)      # fixed sin-cos embedding
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True, epsilon=epsilon) for _ in range(depth)])
        self.norm = LayerNorm(embed_dim, epsilon)
        self.decoder_embed = Linear(embed_dim, decoder_embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, L + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.ModuleList([Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=True, epsilon=epsilon)
                                             for _ in range(decoder_depth)])
        self.decoder_norm = LayerNorm(decoder_embed_dim, epsilon)
        self.decoder_pred = Linear(decoder_embed_dim, patch_size ** 2 * in_chans)
        self.norm_pix_loss = norm_pix_loss
        self.initialize_weights()

    def initialize_weights(self):
        """mae.py:112-154"""
        g = int(self.patch_embed.num_patches ** .5)
        with torch.no_grad():
            self.pos_embed.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.embed_dim, g, True)).float().unsqueeze(0))
            self.decoder_pos_embed.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.decoder_embed_dim, g, True)).float().unsqueeze(0))
        nn.init.normal_(self.cls_token, std=.02)
        nn.init.normal_(self.mask_token, std=.02)
        for lin in (self.decoder_embed, self.decoder_pred):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)

    # ---- reference helper API (host-side, for tests / users) ---------------------------------------------------------
    def patchify(self, imgs):
        p = self.patch_embed.patch_size[0]
        assert imgs.shape[2] == imgs.shape[3] and imgs.shape[2] % p == 0
        h = w = imgs.shape[2] // p
        x = imgs.reshape(imgs.shape[0], 3, h, p, w, p)
        return torch.einsum('nchpwq->nhwpqc', x).reshape(imgs.shape[0], h * w, p * p * 3)

    def random_masking_ids(self, noise, mask_ratio):
        """-> (ids_shuffle, ids_restore, mask, len_keep): mae.py:184-212 with the noise supplied."""
        L = noise.shape[1]
        len_keep = int(L * (1 - mask_ratio))
        s, r, m = V.mae_random_masking(noise.contiguous(), len_keep)
        return s, r, m, len_keep

    # ---- explicit forward / backward ---------------------------------------------------------------------------------
    def _run_forward(self, imgs, noise, mask_ratio, save=True):
        B = imgs.shape[0]
        L, D, Dd = self.patch_embed.num_patches, self.embed_dim, self.decoder_embed_dim
        p = self.patch_embed.patch_size[0]
        imgs = imgs.contiguous()
        pe, cols = self.patch_embed.fwd(imgs, save)
        if noise is None:
            noise = torch.rand(B, L, device=imgs.device)
        ids_shuffle, ids_restore, mask, keep = self.random_masking_ids(noise, mask_ratio)
        # encoder on the visible tokens
        x = V.token_assemble_fwd(pe, self.pos_embed.view(L + 1, D), self.cls_token.view(D), B, L, 1 + keep, D, "mae_enc",
                                 ids=ids_shuffle, keep=keep)
        enc_ctx = []
        for blk in self.blocks:
            x, c = blk.fwd(x, B, 1 + keep, save)
            enc_ctx.append(c)
        latent, cn = self.norm.fwd(x)
        # decoder on the full sequence
        y = K.gemm(latent, compute_copy(self.decoder_embed.weight), bias=self.decoder_embed.bias)
        z = V.token_assemble_fwd(y, self.decoder_pos_embed.view(L + 1, Dd), self.mask_token.view(Dd), B, 1 + keep, L + 1, Dd,
                                 "mae_dec", ids=ids_restore, keep=keep)
        dec_ctx = []
        for blk in self.decoder_blocks:
            z, c = blk.fwd(z, B, L + 1, save)
            dec_ctx.append(c)
        zn, cdn = self.decoder_norm.fwd(z)
        pred_full = K.gemm(zn, compute_copy(self.decoder_pred.weight), bias=self.decoder_pred.bias)     # [B*(L+1), p*p*3]
        mask_sum = float(B * (L - keep))
        Hp = int(L ** .5)
        loss = V.mae_loss_fwd(pred_full, imgs, mask, B, Hp, p, L + 1, 1, self.norm_pix_loss, mask_sum)[0]
        pred = pred_full.view(B, L + 1, -1)[:, 1:, :]                                                    # remove cls token
        saved = (imgs, cols, ids_shuffle, ids_restore, mask, keep, enc_ctx, cn, latent, y, dec_ctx, cdn, zn, pred_full,
                 mask_sum, B) if save else None
        return loss, pred, mask, saved

    def _run_backward(self, saved, dloss):
        (imgs, cols, ids_shuffle, ids_restore, mask, keep, enc_ctx, cn, latent, y, dec_ctx, cdn, zn, pred_full, mask_sum, B) = saved
        L, D, Dd = self.patch_embed.num_patches, self.embed_dim, self.decoder_embed_dim
        p = self.patch_embed.patch_size[0]
        Hp = int(L ** .5)
        dpred = V.mae_loss_bwd(pred_full, imgs, mask, dloss, B, Hp, p, L + 1, 1, self.norm_pix_loss, mask_sum)
        d = _linear_bwd(self.decoder_pred, zn, dpred)
        d = self.decoder_norm.bwd(cdn, d)
        for blk, c in zip(reversed(self.decoder_blocks), reversed(dec_ctx)):
            d = blk.bwd(c, d)
        dy = V.token_assemble_bwd(d, B, 1 + keep, L + 1, Dd, "mae_dec", ids=ids_shuffle, keep=keep,
                                  acc_tok=grad_buffer(self.mask_token).view(Dd), ids_tok=ids_restore)
        d = _linear_bwd(self.decoder_embed, latent, dy)
        d = self.norm.bwd(cn, d)
        for blk, c in zip(reversed(self.blocks), reversed(enc_ctx)):
            d = blk.bwd(c, d)
        dpe = V.token_assemble_bwd(d, B, L, 1 + keep, D, "mae_enc", ids=ids_restore, keep=keep,
                                   acc_tok=grad_buffer(self.cls_token).view(D))
        self.patch_embed.bwd(cols, dpe)

    def forward(self, imgs, mask_ratio=0.75, noise=None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _MAEFn.apply(self, imgs, noise, mask_ratio, self.cls_token)
        loss, pred, mask, _ = self._run_forward(imgs, noise, mask_ratio, save=False)
        return loss, pred, mask


def mae_vit_base_patch16(**kwargs):
    """mae.py:331-344 (mae_vit_base_patch16_dec512d8b)"""
    kw = dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, decoder_embed_dim=512, decoder_depth=8,
              decoder_num_heads=16, mlp_ratio=4, epsilon=1e-6)
    kw.update(kwargs)
    return MaskedAutoencoderViT(**kw)
