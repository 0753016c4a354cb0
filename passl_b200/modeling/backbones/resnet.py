"""ResNet backbone (v1.5 bottleneck: stride on the 3x3) on the tcgen05 implicit-GEMM kernels, NHWC bf16.

Mirrors passl_v110/modeling/backbones/resnet.py:25-77 (`ResNet(depth, num_classes=0, with_pool=False,
zero_init_residual=False, frozen_stages=-1)`, kaiming fan_out init :75-88) over the structure of
resnetimagenet.py:93-246, and `ResNetsimclr` (resnetsimclr.py:25-91 over resnetcifar.py:216-334: same net without the
stem max-pool).  Input is the reference's NCHW fp32 image batch; output is NHWC bf16 [B, h, w, 2048]
(or [B, 2048] with with_pool).

The whole backbone is ONE autograd node: forward chains the fused conv+BN(+ReLU/+residual) units, backward walks them
in reverse with hand-written dgrad / wgrad / BN-backward kernels and accumulates parameter gradients in place.
"""
import math

import torch
import torch.nn as nn

from ... import kernels as K
from ...core.streams import join
from ...core.param_store import grad_buffer
from ...nn.layers import BatchNormState, ConvBN, _kaiming_normal_fan_out
from ..registry import BACKBONES

STEM_K = 7 * 7 * 3
STEM_KPAD = 152  # multiple of 8 (16-byte rows for TMA); the tail is zero


class Stem(nn.Module):
    """conv 7x7/2 (3->64) as repack + 4x1 implicit-GEMM tcgen05 conv (csrc/stem.cu), BN, ReLU, optional 3x3/2 max-pool."""

    def __init__(self, maxpool=True):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(64, STEM_KPAD))      # [Cout, (r, s, c) padded]
        w = torch.empty(64, 7, 7, 3)
        _kaiming_normal_fan_out(w, 64 * 7 * 7)
        with torch.no_grad():
            self.weight[:, :STEM_K] = w.reshape(64, STEM_K)
        self.bn = BatchNormState(64)
        self.maxpool = maxpool

    def fwd(self, img, training=True, save=True):
        N, _, H, W = img.shape
        Ho, Wo = H // 2, W // 2
        # no im2col matrix: W-unfolded space-to-depth repack (1.6 GB / 1024 images instead of 3.9 GB) + a 4x1 implicit-GEMM conv
        cols = K.stem_pack_input(img)
        bn = self.bn
        batch_stats = training and not bn.use_global_stats
        w = K.stem_pack_weight(self.weight.detach())
        if batch_stats:
            stats = K.stats_buffer(64, img.device)
            y = K.stem_conv_fwd(cols, w, col_stats=stats).view(N * Ho * Wo, 64)
            msss = K.bn_finalize(stats, bn.weight, bn.bias, bn._mean, bn._variance, y.shape[0], eps=bn.eps, momentum=bn.momentum)
        else:
            y = K.stem_conv_fwd(cols, w).view(N * Ho * Wo, 64)
            msss = bn.global_affine()
        if self.maxpool:
            # BN apply + ReLU + max-pool in one kernel: the normalised 112^2 activation is never written (the backward
            # recomputes the ReLU mask from y and only needs the pool's arg-max)
            z = None
            out, arg = K.bn_relu_maxpool_fwd(y.view(N, Ho, Wo, 64), msss)
        else:
            z = K.bn_apply(y, msss, True).view(N, Ho, Wo, 64)
            out, arg = z, None
        return out, ((cols, y, z, msss, arg) if save else None)

    def bwd(self, ctx, dout):
        cols, y, z, msss, arg = ctx
        N, Ho, Wo, _ = cols.shape
        dz = K.maxpool_bwd(dout, arg, (N, Ho, Wo, 64)) if arg is not None else dout
        bn = self.bn
        tb = bn.weight.requires_grad
        dy, _, _ = K.bn_bwd(y, dz.view(y.shape), None if z is None else z.view(y.shape), msss, bn.weight, True,
                            dgamma=grad_buffer(bn.weight) if tb else None, dbeta=grad_buffer(bn.bias) if tb else None)
        if self.weight.requires_grad:
            K.stem_conv_wgrad(cols, dy.view(cols.shape[0], cols.shape[1], cols.shape[2], 64), grad_buffer(self.weight))


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=False):
        super().__init__()
        self.conv1 = ConvBN(inplanes, planes, 1, relu=True)
        self.conv2 = ConvBN(planes, planes, 3, stride=stride, pad=1, relu=True)
        self.conv3 = ConvBN(planes, planes * 4, 1, relu=True)          # ReLU applied after the residual add
        self.downsample = ConvBN(inplanes, planes * 4, 1, stride=stride, relu=False) if downsample else None

    def fwd(self, x, training=True, save=True):
        o1, c1 = self.conv1.fwd(x, training=training, save=save)
        o2, c2 = self.conv2.fwd(o1, training=training, save=save)
        if self.downsample is not None:
            idn, cd = self.downsample.fwd(x, training=training, save=save)
        else:
            idn, cd = x, None
        o3, c3 = self.conv3.fwd(o2, residual=idn, training=training, save=save)
        return o3, ((c1, c2, c3, cd) if save else None)

    def bwd(self, ctx, dout, need_dx=True):
        c1, c2, c3, cd = ctx
        d_o2, d_idn = self.conv3.bwd(c3, dout)
        d_o1, _ = self.conv2.bwd(c2, d_o2)
        if self.downsample is not None:
            dx, _ = self.downsample.bwd(cd, d_idn, need_dx=need_dx)
            self.conv1.bwd(c1, d_o1, need_dx=need_dx, dx_out=dx, accumulate=True)
        else:
            dx = d_idn
            self.conv1.bwd(c1, d_o1, need_dx=True, dx_out=dx, accumulate=True)
        return dx


class _BackboneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, img, anchor):
        out, saved = module._run_forward(img, training=module.training, save=torch.is_grad_enabled() or True)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.module._run_backward(ctx.saved, dout.contiguous())
        ctx.saved = None
        return None, None, None


@BACKBONES.register()
class ResNet(nn.Module):
    LAYER_CFG = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}

    def __init__(self, depth=50, num_classes=0, with_pool=False, zero_init_residual=False, frozen_stages=-1,
                 pretrained=None, stem_maxpool=True):
        super().__init__()
        if depth not in self.LAYER_CFG:
            raise ValueError("passl_b200 ResNet supports bottleneck depths %s" % sorted(self.LAYER_CFG))
        assert num_classes <= 0, "classification fc is outside the self-supervised hot path"
        self.with_pool = with_pool
        self._layers = tuple(self.LAYER_CFG[depth])     # block counts per stage (checkpoint name mapping reads it)
        self.stem = Stem(maxpool=stem_maxpool)
        blocks, inplanes = [], 64
        for i, (planes, n) in enumerate(zip([64, 128, 256, 512], self.LAYER_CFG[depth])):
            stride = 1 if i == 0 else 2
            for b in range(n):
                s = stride if b == 0 else 1
                blocks.append(Bottleneck(inplanes, planes, s, downsample=(b == 0 and (s != 1 or inplanes != planes * 4))))
                inplanes = planes * 4
        self.blocks = nn.ModuleList(blocks)
        self.out_channels = inplanes
        if zero_init_residual:          # passl/models/resnet.py:68-73, resnet.py(v110):81-86
            for blk in self.blocks:
                nn.init.zeros_(blk.conv3.bn.weight)
        if frozen_stages >= 0:          # resnet.py(v110):90-105 — linear-probe / detection fine-tuning, not the pre-training path
            raise NotImplementedError("frozen_stages >= 0 (frozen stem / stages with global-statistics BatchNorm) is not built: "
                                      "the explicit backward here runs through every stage")
        if pretrained is not None:      # resnet.py(v110):63-72: a .pdparams state dict, optionally wrapped in {'state_dict': ...}
            from ...utils import checkpoint as C
            state = C.load_pdparams(pretrained)
            C.resnet_from_paddle(self, state["state_dict"] if isinstance(state.get("state_dict"), dict) else state)

    # -- explicit forward / backward -------------------------------------------------------------------------
    def _run_forward(self, img, training=True, save=True):
        assert img.dim() == 4 and img.shape[1] == 3 and img.dtype == torch.float32, "expects NCHW fp32 images"
        x, cs = self.stem.fwd(img.contiguous(), training=training, save=save)
        ctxs = []
        for blk in self.blocks:
            x, c = blk.fwd(x, training=training, save=save)
            ctxs.append(c)
        pooled = None
        if self.with_pool:
            pooled, _ = K.avgpool_fwd(x)
        return (pooled if self.with_pool else x), (cs, ctxs, tuple(x.shape))

    def _run_backward(self, saved, dout):
        cs, ctxs, feat_shape = saved
        d = K.avgpool_bwd(dout, feat_shape) if self.with_pool else dout
        for blk, c in zip(reversed(self.blocks), reversed(ctxs)):
            d = blk.bwd(c, d)
        self.stem.bwd(cs, d)
        join()          # the weight-gradient GEMMs ran on the side stream (core/streams.py)

    def forward(self, img):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _BackboneFn.apply(self, img, self.stem.weight)
        out, _ = self._run_forward(img, training=self.training, save=False)
        return out


@BACKBONES.register()
class ResNetsimclr(ResNet):
    """SimCLR backbone (resnetsimclr.py:25-91): the stem max-pool is absent (resnetcifar.py:275,321-332)."""

    def __init__(self, depth=50, **kw):
        kw.setdefault("stem_maxpool", False)
        kw.setdefault("with_pool", True)
        super().__init__(depth=depth, **kw)
        # every Conv2D of resnetcifar.py carries XavierNormal(fan_in=None, fan_out=0) (resnetcifar.py:137-141,261-264) and the
        # subclass leaves init_parameters() commented out (resnetsimclr.py:62): std = sqrt(2 / (fan_in + 0)), fan_in = Cin*k*k —
        # not the fan_out rule of the MoCo backbone
        with torch.no_grad():
            self.stem.weight.zero_()
            self.stem.weight[:, :STEM_K].normal_(0.0, math.sqrt(2.0 / STEM_K))
            for m in self.modules():
                if isinstance(m, ConvBN):
                    m.weight.normal_(0.0, math.sqrt(2.0 / (m.cin * m.k * m.k)))
