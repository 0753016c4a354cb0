"""Projection necks of the contrastive models (passl_v110/modeling/necks/base_neck.py).

  LinearNeck        :43-64    avgpool -> fc
  NonLinearNeckV1   :67-94    avgpool -> fc -> ReLU -> fc                      (MoCo v2: 2048 -> 2048 -> 128)
  NonLinearNeckfc3  :209-237  fc-BN1D-ReLU, fc-BN1D-ReLU, fc-BN1D, l2_normalize (SimCLR)

Input: backbone features, NHWC bf16 [B, h, w, C] (pooled here when with_avg_pool) or [B, C].
Output: fp32 [B, out_channels] embeddings (the heads / fused InfoNCE kernels take fp32 + a bf16 copy).
Each neck is one autograd node with a hand-written backward over the tcgen05 GEMM (dgrad / wgrad) kernels.
"""
import torch
import torch.nn as nn

from ... import kernels as K
from ...nn.layers import BatchNorm1D, Linear
from ..registry import NECKS


class _NeckFn(torch.autograd.Function):
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


class _NeckBase(nn.Module):
    def _pool(self, x):
        if x.dim() == 4:
            if self.with_avg_pool:
                pooled, _ = K.avgpool_fwd(x)
                return pooled, tuple(x.shape)
            assert x.shape[1] == 1 and x.shape[2] == 1, "un-pooled feature map needs with_avg_pool=True"
            return x.reshape(x.shape[0], -1), None
        return x, None

    def _unpool(self, dx2d, feat_shape):
        return K.avgpool_bwd(dx2d, feat_shape) if feat_shape is not None else dx2d

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _NeckFn.apply(self, x, next(self.parameters()))
        out, _ = self._run_forward(x, training=self.training, save=False)
        return out


@NECKS.register()
class LinearNeck(_NeckBase):
    def __init__(self, in_channels, out_channels, with_avg_pool=True):
        super().__init__()
        self.with_avg_pool = with_avg_pool
        self.fc = Linear(in_channels, out_channels, out_fp32=True)

    def _run_forward(self, x, training=True, save=True):
        h, feat_shape = self._pool(x)
        out, c = self.fc.fwd(h, save=save)
        return out, (c, feat_shape, x.dim())

    def _run_backward(self, saved, dout):
        c, feat_shape, _ = saved
        dx = self.fc.bwd(c, K.cast_bf16(dout))
        return self._unpool(dx, feat_shape)


@NECKS.register()
class NonLinearNeckV1(_NeckBase):
    """fc-relu-fc (MoCo v2)."""

    def __init__(self, in_channels, hid_channels, out_channels, with_avg_pool=True):
        super().__init__()
        self.with_avg_pool = with_avg_pool
        self.fc1 = Linear(in_channels, hid_channels, act="relu")
        self.fc2 = Linear(hid_channels, out_channels, out_fp32=True)
        for fc in (self.fc1, self.fc2):   # _init_parameters(init_linear='kaiming'): kaiming fan_in, relu; bias 0
            nn.init.kaiming_normal_(fc.weight, mode="fan_in", nonlinearity="relu")

    def _run_forward(self, x, training=True, save=True):
        h0, feat_shape = self._pool(x)
        h1, c1 = self.fc1.fwd(h0, save=save)
        out, c2 = self.fc2.fwd(h1, save=save)
        return out, (c1, c2, feat_shape)

    def _run_backward(self, saved, dout):
        c1, c2, feat_shape = saved
        h1 = c2[0]
        dh1 = self.fc2.bwd(c2, K.cast_bf16(dout), prev_act_out=h1, prev_act="relu_mask")   # ReLU mask fused in the epilogue
        dx = self.fc1.bwd(c1, dh1)
        return self._unpool(dx, feat_shape)


@NECKS.register()
class NonLinearNeckfc3(_NeckBase):
    """fc-BN-ReLU, fc-BN-ReLU, fc-BN, then l2_normalize(hidden, -1) (base_neck.py:231-237)."""

    def __init__(self, in_channels, hid_channels, out_channels, with_avg_pool=True):
        super().__init__()
        self.with_avg_pool = with_avg_pool
        self.fc1, self.bn1 = Linear(in_channels, hid_channels), BatchNorm1D(hid_channels, relu=True)
        self.fc2, self.bn2 = Linear(hid_channels, hid_channels), BatchNorm1D(hid_channels, relu=True)
        self.fc3, self.bn3 = Linear(hid_channels, out_channels), BatchNorm1D(out_channels, relu=False)
        for fc in (self.fc1, self.fc2, self.fc3):     # init_backbone_weight_simclr: normal(0, 0.01), bias 0
            nn.init.normal_(fc.weight, 0.0, 0.01)

    def _run_forward(self, x, training=True, save=True):
        h0, feat_shape = self._pool(x)
        dev = h0.device
        ctxs = []
        h = h0
        for i, (fc, bn) in enumerate(((self.fc1, self.bn1), (self.fc2, self.bn2), (self.fc3, self.bn3))):
            y, cf = fc.fwd(h, save=save)
            h, cb = bn.fwd(y, training=training, save=save, out_f32=(i == 2))
            ctxs.append((cf, cb))
        emb, _, inv = K.l2norm_fwd(h, mode="l2_normalize")
        return emb, (ctxs, feat_shape, emb, inv)

    def _run_backward(self, saved, dout):
        ctxs, feat_shape, emb, inv = saved
        _, d = K.l2norm_bwd(dout, emb, inv, mode="l2_normalize", want_bf16=True)
        for (fc, bn), (cf, cb) in zip(((self.fc3, self.bn3), (self.fc2, self.bn2), (self.fc1, self.bn1)), reversed(ctxs)):
            dy = bn.bwd(cb, d)
            d = fc.bwd(cf, dy)
        return self._unpool(d, feat_shape)
