"""MoCo v3 (passl/models/mocov3.py:37-297; cross-check tasks/ssl/mocov3/builder_moco.py:18-159) on the B200 kernels.

  base_encoder   = ViT (fixed 2-D sin-cos position embedding, mocov3.py:58-91) + 3-layer projector MLP (Linear(no bias)-BN1D-ReLU x2,
                   Linear-BN1D(no affine)), predictor = 2-layer MLP of the same kind (mocov3.py:136-169)
  momentum_encoder = EMA copy of base_encoder (projector included), cosine momentum schedule (builder_moco.py:74-80)
  loss           = ctr(q1, k2) + ctr(q2, k1),  ctr: normalize -> all_gather(k) (no grad) -> q.k^T/T -> labels arange(N)+N*rank
                   -> CE * 2T   (mocov3.py:187-198)   — computed by the fused tcgen05 InfoNCE kernel (label mode).
The reference wraps `nn.Sequential(base_encoder, predictor)` in its CosineEMA with momentum weighting the *source*
(mocov3.py:133-134, averaged_model.py:165-186); set `reference_ema_quirk=True` to reproduce that formula.
"""
import math

import torch
import torch.nn as nn

from .. import kernels as K
from ..core.param_store import ParamStore
from ..distributed import concat_all_gather, get_rank, get_world_size
from ..loss.contrastive import gathered_infonce, normalize
from ..nn.layers import BatchNorm1D, Linear
from .vision_transformer import VisionTransformer


def mocov3_sincos_pos_embed(embed_dim, grid_h, grid_w, temperature=10000.0):
    """Fixed position table of MoCo v3 (mocov3.py:67-91), [1, 1 + grid_h*grid_w, embed_dim] fp32, class-token row zero.

    Row p of the table is [sin(a w), cos(a w), sin(b w), cos(b w)] with frequencies w_i = temperature^(-i / (D/4)) and
    (a, b) = (p // grid_h, p % grid_h): the reference meshgrids (arange(w), arange(h)) with 'ij' indexing and flattens, so the
    slow coordinate comes first.  This differs from the MAE table (models/mae.py) by the order of the two halves."""
    assert embed_dim % 4 == 0, "Embed dimension must be divisible by 4 for 2D sin-cos position embedding"
    pos_dim = embed_dim // 4
    omega = 1.0 / (temperature ** (torch.arange(pos_dim, dtype=torch.float64) / pos_dim))
    p = torch.arange(grid_h * grid_w, dtype=torch.float64)
    a, b = torch.div(p, grid_h, rounding_mode="floor"), torch.remainder(p, grid_h)
    oa, ob = a[:, None] * omega[None], b[:, None] * omega[None]
    table = torch.cat([oa.sin(), oa.cos(), ob.sin(), ob.cos()], dim=1)
    return torch.cat([torch.zeros(1, embed_dim, dtype=torch.float64), table], dim=0).unsqueeze(0).float()


class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, anchor):
        out, saved = module._run_forward(x, training=module.training, save=True)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        dx = ctx.module._run_backward(ctx.saved, dout.contiguous())
        ctx.saved = None
        return None, dx, None


class MLPBN(nn.Module):
    """_build_mlp (mocov3.py:136-158): Linear(bias=False) -> BN1D -> ReLU ... -> Linear(bias=False) -> BN1D(no affine).
    bf16 [B, in] -> fp32 [B, out]."""

    def __init__(self, num_layers, input_dim, mlp_dim, output_dim, last_bn=True):
        super().__init__()
        fcs, bns = [], []
        for l in range(num_layers):
            d1 = input_dim if l == 0 else mlp_dim
            d2 = output_dim if l == num_layers - 1 else mlp_dim
            fcs.append(Linear(d1, d2, bias=False))
            if l < num_layers - 1:
                bns.append(BatchNorm1D(d2, relu=True))
            else:
                bns.append(BatchNorm1D(d2, relu=False, affine=False) if last_bn else None)
        self.fcs, self.bns = nn.ModuleList(fcs), nn.ModuleList([b for b in bns if b is not None])
        for fc in self.fcs:                              # paddle nn.Linear default: Xavier-uniform weight (bias_attr=False here)
            nn.init.xavier_uniform_(fc.weight)
        self.last_bn = last_bn

    def _run_forward(self, x, training=True, save=True):
        if x.dtype != torch.bfloat16:
            x = K.cast_bf16(x.contiguous())
        ctxs, h, n = [], x, len(self.fcs)
        for i, fc in enumerate(self.fcs):
            y, cf = fc.fwd(h, save=save)
            if i < len(self.bns):
                h, cb = self.bns[i].fwd(y, training=training, save=save, out_f32=(i == n - 1))
            else:
                h, cb = K.cast_f32(y), None
            ctxs.append((cf, cb))
        return h, ctxs

    def _run_backward(self, ctxs, dout):
        d = K.cast_bf16(dout)
        for i in reversed(range(len(self.fcs))):
            cf, cb = ctxs[i]
            if cb is not None:
                d = self.bns[i].bwd(cb, d)
            d = self.fcs[i].bwd(cf, d)
        return d

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _MLPFn.apply(self, x, self.fcs[0].weight)
        out, _ = self._run_forward(x, training=self.training, save=False)
        return out


class MoCoV3ViT(VisionTransformer):
    """ViT with the fixed 2-D sin-cos position embedding (mocov3.py:37-91)."""

    def __init__(self, stop_grad_conv1=False, **kwargs):
        kwargs.setdefault("learnable_pos", False)
        super().__init__(**kwargs)
        # weight initialisation of mocov3.py:43-61: q, k, v treated as three separate Xavier-uniform matrices, every other Linear
        # Xavier-uniform with zero bias (the Block constructor already does that), cls_token ~ N(0, 1e-6), patch projection
        # uniform with the fan of a [3 p p] -> [D] matrix
        with torch.no_grad():
            D = self.embed_dim
            for blk in self.blocks:
                val = math.sqrt(6.0 / float(D + D))              # weight.shape[1] // 3 + weight.shape[0] of the [in, 3 out] matrix
                blk.qkv.weight.uniform_(-val, val)
            self.cls_token.normal_(std=1e-6)
            pw = self.patch_embed.proj.weight
            val = math.sqrt(6.0 / float(pw.shape[1] + D))        # 3 * prod(patch_size) + embed_dim
            pw.uniform_(-val, val)
            self.patch_embed.proj.bias.zero_()
        if stop_grad_conv1:                          # mocov3.py:63-65: the patch projection stays at its random initialisation
            self.patch_embed.proj.weight.requires_grad = False
            self.patch_embed.proj.bias.requires_grad = False
        g = int(self.patch_embed.num_patches ** .5)
        with torch.no_grad():
            self.pos_embed.copy_(mocov3_sincos_pos_embed(self.embed_dim, g, g))
        self.pos_embed.requires_grad = False


class _Encoder(nn.Module):
    def __init__(self, vit, head):
        super().__init__()
        self.vit, self.head = vit, head

    def forward(self, x):
        return self.head(self.vit(x))


class MoCoV3Pretrain(nn.Module):
    def __init__(self, base_encoder, dim=256, mlp_dim=4096, T=1.0, base_momentum=0.99, max_steps=1000,
                 reference_ema_quirk=False, peer_loss=None):
        super().__init__()
        self.T, self.base_momentum, self.max_steps, self.quirk = T, base_momentum, max_steps, reference_ema_quirk
        # fused compute + collective (single node): the loss kernels read every rank's key shard in place over NVLink instead of
        # consuming an NCCL all-gathered copy (distributed/peer.py::PeerKeyShards); PASSL_B200_PEER_LOSS=1 or peer_loss=True
        import os
        self.peer_loss = bool(int(os.environ.get("PASSL_B200_PEER_LOSS", "0"))) if peer_loss is None else bool(peer_loss)
        self._shards = None
        vit_q, vit_k = base_encoder(), base_encoder()
        hidden = vit_q.embed_dim
        self.base_encoder = _Encoder(vit_q, MLPBN(3, hidden, mlp_dim, dim))
        self.predictor = MLPBN(2, dim, mlp_dim, dim)
        self.momentum_encoder = _Encoder(vit_k, MLPBN(3, hidden, mlp_dim, dim))
        # The reference wraps Sequential(base_encoder, predictor) in CosineEMA and calls THAT for the keys (mocov3.py:133-134,
        # 224-225; averaged_model.py:59-61), so its targets also pass through an averaged copy of the predictor.  Literal mode
        # (reference_ema_quirk) keeps that copy; the default is the textbook momentum encoder of builder_moco.py:36-60.
        self.momentum_predictor = MLPBN(2, dim, mlp_dim, dim) if reference_ema_quirk else None
        pairs = [(self.base_encoder, self.momentum_encoder)] + ([(self.predictor, self.momentum_predictor)] if reference_ema_quirk else [])
        for src, dst in pairs:
            for pq, pk in zip(src.parameters(), dst.parameters()):
                pk.data.copy_(pq.data)
                pk.requires_grad = False
        self.steps = 0
        self._stores = None

    def build_param_stores(self):
        """flat storage: (base_encoder + predictor) trainable, momentum encoder (+ its predictor copy in literal mode) EMA target"""
        # containers only for enumeration: kept out of the module tree so state_dict() lists every tensor once
        self.__dict__["_trainable"] = nn.ModuleList([self.base_encoder, self.predictor])
        st = ParamStore(self._trainable, with_grad=True)
        self.__dict__["_averaged"] = nn.ModuleList([self.momentum_encoder] + ([self.momentum_predictor] if self.quirk else []))
        sk = ParamStore(self._averaged, with_grad=False)
        # EMA runs over a prefix of the trainable buffer: both stores enumerate their tensors in the same order
        self._ema_numel = sk.numel
        self._stores = (st, sk)
        return st, sk

    def _keys(self, x):
        k = self.momentum_encoder(x)
        return self.momentum_predictor(k) if self.quirk else k

    @torch.no_grad()
    def _update_momentum_encoder(self):
        st, sk = self._stores
        if self.quirk:      # averaged*(1-m') + source*m' with m' cosine from base to end 0 (averaged_model.py:165-186)
            m_src = self.base_momentum * (math.cos(math.pi * self.steps / float(self.max_steps)) + 1) / 2
            m = 1.0 - m_src
        else:               # builder_moco.py:74-80 / main_moco.py adjust_moco_momentum
            m = 1. - 0.5 * (1. + math.cos(math.pi * self.steps / float(self.max_steps))) * (1. - self.base_momentum)
        K.ema_update(sk.master, st.master[:self._ema_numel], m, k_bf16=sk.bf16)
        self.steps += 1

    def contrastive_loss(self, q, k):
        q = normalize(q)
        if self.peer_loss and get_world_size() > 1 and q.shape[0] % 64 == 0 and q.shape[1] % 64 == 0 and q.shape[1] <= 256:
            from ..distributed.peer import PeerKeyShards, peer_gathered_infonce
            if self._shards is None or (self._shards.n, self._shards.d) != tuple(q.shape):
                if self._shards is not None:
                    self._shards.close()
                self._shards = PeerKeyShards(q.shape[0], q.shape[1])
            with torch.no_grad():
                kn = normalize(k)
            loss, _, _ = peer_gathered_infonce(q, kn, self._shards, 1.0 / self.T, 2 * self.T)
            return loss
        with torch.no_grad():
            k = concat_all_gather(normalize(k))
        N = q.shape[0]
        labels = torch.arange(N, dtype=torch.int64, device=q.device) + N * get_rank()
        loss, _, _ = gathered_infonce(q, k, labels, 1.0 / self.T, loss_scale=2 * self.T)
        return loss

    def forward(self, inputs):
        assert isinstance(inputs, list)
        x1, x2 = inputs[0], inputs[1]
        if self._stores is None:
            self.build_param_stores()
        q1 = self.predictor(self.base_encoder(x1))
        q2 = self.predictor(self.base_encoder(x2))
        with torch.no_grad():
            self._update_momentum_encoder()
            k1 = self._keys(x1)
            k2 = self._keys(x2)
        return _Add.apply(self.contrastive_loss(q1, k2), self.contrastive_loss(q2, k1))


class _Add(torch.autograd.Function):
    """loss_a + loss_b on device scalars through the axpy kernel (no torch math on the path)."""

    @staticmethod
    def forward(ctx, a, b):
        out = a.detach().clone().reshape(1)
        K.axpy(out, b.detach().reshape(1).contiguous(), 1.0)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        return g, g


def mocov3_vit_base_pretrain(**kwargs):
    """mocov3.py:288-297"""
    def enc():
        return MoCoV3ViT(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True, epsilon=1e-6,
                         stop_grad_conv1=True)       # mocov3.py:289
    kw = dict(dim=256, mlp_dim=4096, T=0.2)
    kw.update(kwargs)
    return MoCoV3Pretrain(enc, **kw)
