"""Building blocks with explicit forward / backward over the C-ABI kernels (channels-last bf16 activations).

Every unit exposes ``fwd(x, ...) -> (out, ctx)`` and ``bwd(ctx, dout, ...) -> dx``; parameter gradients are
accumulated in place into the fp32 gradient buffer of the parameter (ParamStore view), the way the reference's
fused-parameter storage accumulates (passl/core/param_fuse.py).  torch is only the tensor carrier: no torch math
runs on the hot path.

Reference layers replaced: paddle nn.Conv2D + nn.BatchNorm2D + nn.ReLU (resnetimagenet.py:112-148),
nn.Linear / nn.BatchNorm1D (necks/base_neck.py), nn.MaxPool2D, nn.AdaptiveAvgPool2D.
"""
import math

import torch
import torch.nn as nn

from .. import kernels as K
from ..core.param_store import compute_copy, grad_buffer
from ..core.streams import side_stream


def _kaiming_normal_fan_out(w, fan_out):
    """passl_v110/modules/init.py kaiming_init(mode='fan_out', nonlinearity='relu'): std = sqrt(2 / fan_out)."""
    with torch.no_grad():
        w.normal_(0.0, math.sqrt(2.0 / fan_out))


class BatchNormState(nn.Module):
    """gamma / beta + running statistics of one BatchNorm layer (Paddle: eps 1e-5, momentum 0.9)."""

    def __init__(self, c, eps=1e-5, momentum=0.9, affine=True):
        super().__init__()
        self.c, self.eps, self.momentum = c, eps, momentum
        if affine:
            self.weight = nn.Parameter(torch.ones(c))
            self.bias = nn.Parameter(torch.zeros(c))
        else:
            self.weight = self.bias = None
        self.register_buffer("_mean", torch.zeros(c))
        self.register_buffer("_variance", torch.ones(c))
        self.use_global_stats = False  # passl_v110/modules/freeze.py:17-23

    def global_affine(self):
        """[4, C] (mean, invstd, scale, shift) from the running statistics (use_global_stats / eval mode)."""
        return K.bn_global_affine(self._mean, self._variance, self.weight, self.bias, eps=self.eps)


class ConvBN(nn.Module):
    """Conv2D(bias=False) -> BatchNorm2D -> (+residual) -> (ReLU), NHWC bf16, weights [Cout, R, S, Cin]."""

    def __init__(self, cin, cout, k, stride=1, pad=0, relu=True):
        super().__init__()
        self.cin, self.cout, self.k, self.stride, self.pad, self.relu = cin, cout, k, stride, pad, relu
        self.weight = nn.Parameter(torch.empty(cout, k, k, cin))
        _kaiming_normal_fan_out(self.weight, cout * k * k)
        self.bn = BatchNormState(cout)

    def fwd(self, x, residual=None, training=True, save=True):
        w = compute_copy(self.weight)
        bn = self.bn
        batch_stats = training and not bn.use_global_stats
        if batch_stats:
            # batch statistics come out of the conv epilogue (per-CTA partials of the bf16 values it stores): no extra pass over y
            stats = K.stats_buffer(self.cout, x.device)
            y = K.conv2d_fwd(x, w, stride=self.stride, pad=self.pad, col_stats=stats)
            count = y.numel() // self.cout
            msss = K.bn_finalize(stats, bn.weight, bn.bias, bn._mean, bn._variance, count, eps=bn.eps, momentum=bn.momentum)
        else:
            y = K.conv2d_fwd(x, w, stride=self.stride, pad=self.pad)
            msss = bn.global_affine()
        mask = None
        if residual is not None and self.relu and save:
            # residual units: the backward needs relu'(.) of the SUM — keep it as 1 bit per element instead of re-reading z
            z, mask = K.bn_apply_mask(y, msss, residual=residual)
        else:
            z = K.bn_apply(y, msss, self.relu, residual=residual)
        ctx = (x, y, z if mask is None else None, msss, residual is not None, batch_stats, mask) if save else None
        return z, ctx

    def bwd(self, ctx, dz, need_dx=True, dx_out=None, accumulate=False):
        """Returns (dx, dres).  dx is written into dx_out (accumulated when accumulate) if given."""
        x, y, z, msss, has_res, batch_stats, mask = ctx
        bn = self.bn
        assert batch_stats, "backward through a use_global_stats BatchNorm is not on the training path"
        train_bn = bn.weight is not None and bn.weight.requires_grad
        dy, dres, _ = K.bn_bwd(y, dz, z, msss, bn.weight, self.relu, want_dres=has_res,
                               dgamma=grad_buffer(bn.weight) if train_bn else None,
                               dbeta=grad_buffer(bn.bias) if train_bn else None, mask_bits=mask)
        if self.weight.requires_grad:
            with side_stream(x, dy):      # off the dgrad critical path: overlaps the next unit's HBM-bound BN backward
                K.conv2d_wgrad(x, dy, tuple(self.weight.shape), stride=self.stride, pad=self.pad,
                               out=grad_buffer(self.weight), accumulate=True)
        dx = None
        if need_dx:
            dx = K.conv2d_dgrad(dy, compute_copy(self.weight), tuple(x.shape), stride=self.stride, pad=self.pad,
                                out=dx_out, accumulate=accumulate)
        return dx, dres


class Linear(nn.Module):
    """y = act(x W^T + b); W stored [out, in] (Paddle stores [in, out]; transposed at checkpoint I/O)."""

    def __init__(self, cin, cout, bias=True, act=None, out_fp32=False):
        super().__init__()
        self.cin, self.cout, self.act, self.out_fp32 = cin, cout, act, out_fp32
        self.weight = nn.Parameter(torch.empty(cout, cin))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None

    def fwd(self, x, save=True, col_stats=None):
        w = compute_copy(self.weight)
        out = K.gemm(x, w, bias=self.bias, act=self.act, out_dtype=torch.float32 if self.out_fp32 else torch.bfloat16,
                     col_stats=col_stats)
        return out, ((x, out) if save else None)

    def bwd(self, ctx, dout_bf16, need_dx=True, prev_act_out=None, prev_act="relu_mask"):
        """dout_bf16 is the gradient w.r.t. the pre-activation output of this layer ([M, out] bf16).
        prev_act_out: output of a ReLU that produced x -> the dx GEMM applies the mask in its epilogue."""
        x, _ = ctx
        M = x.shape[0]
        if self.weight.requires_grad:
            K.gemm(dout_bf16, x, a_t=True, b_t=True, out=grad_buffer(self.weight), accumulate=True,
                   splits=K.wgrad_splits(self.cout, self.cin, M))
            if self.bias is not None:
                K.colsum_accumulate(dout_bf16, grad_buffer(self.bias))
        if not need_dx:
            return None
        return K.gemm(dout_bf16, compute_copy(self.weight), b_t=True, aux=prev_act_out, aux_mode_name=prev_act)


class BatchNorm1D(nn.Module):
    """BatchNorm over [B, C] bf16 features (necks/base_neck.py:221-227), optional fused ReLU."""

    def __init__(self, c, relu=False, affine=True):
        super().__init__()
        self.c, self.relu = c, relu
        self.bn = BatchNormState(c, affine=affine)

    def fwd(self, y, stats=None, training=True, save=True, out_f32=False):
        bn = self.bn
        batch_stats = training and not bn.use_global_stats
        if batch_stats:
            if stats is None:
                stats = K.bn_stats(y)
            msss = K.bn_finalize(stats, bn.weight, bn.bias, bn._mean, bn._variance, y.shape[0], eps=bn.eps, momentum=bn.momentum)
        else:
            msss = bn.global_affine()
        if out_f32:
            zf = torch.empty(y.shape, dtype=torch.float32, device=y.device)
            z = K.bn_apply(y, msss, self.relu, out=torch.empty_like(y), out_f32=zf)
        else:
            zf = None
            z = K.bn_apply(y, msss, self.relu)
        return (zf if out_f32 else z), ((y, z, msss) if save else None)

    def bwd(self, ctx, dz_bf16):
        y, z, msss = ctx
        bn = self.bn
        train_bn = bn.weight is not None and bn.weight.requires_grad
        dy, _, _ = K.bn_bwd(y, dz_bf16, z, msss, bn.weight, self.relu,
                            dgamma=grad_buffer(bn.weight) if train_bn else None,
                            dbeta=grad_buffer(bn.bias) if train_bn else None)
        return dy
