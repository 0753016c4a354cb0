"""Where does the backward of one Bottleneck leave the quantisation-matched oracle?  Every intermediate gradient of the CUDA unit
(dy of each BatchNorm backward, the dgrad outputs, the weight gradients) against the oracle's gradient at the same tensor.
Developer diagnostic (round 2): python tools/diag_bwd.py [inplanes planes stride ds B HW]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import resnet as O  # noqa: E402  (developer diagnostic: same role as tests/)
from passl_b200 import kernels as K  # noqa: E402
from passl_b200.core.param_store import compute_copy  # noqa: E402
from passl_b200.modeling.backbones.resnet import Bottleneck  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nchw(t):
    return t.float().cpu().double().permute(0, 3, 1, 2).contiguous()


def main(inpl=1024, planes=256, stride=1, ds=0, B=16, HW=8):
    torch.manual_seed(0)
    blk = Bottleneck(inpl, planes, stride, downsample=bool(ds)).cuda()
    x = torch.randn(B, HW, HW, inpl, device="cuda").relu().bfloat16()
    out, (c1, c2, c3, cd) = blk.fwd(x)
    dout = torch.randn_like(out)
    for p in blk.parameters():
        p.grad = torch.zeros_like(p)
    # ---- CUDA backward, step by step (ConvBN.bwd opened up) ----
    def unit_bwd(unit, ctx, dz, tag, rec):
        xin, y, z, msss, has_res, _, mask = ctx
        dy, dres, sums = K.bn_bwd(y, dz, z, msss, unit.bn.weight, unit.relu, want_dres=has_res, mask_bits=mask)
        dx = K.conv2d_dgrad(dy, compute_copy(unit.weight), tuple(xin.shape), stride=unit.stride, pad=unit.pad)
        dw = torch.zeros_like(unit.weight)
        K.conv2d_wgrad(xin, dy, tuple(unit.weight.shape), stride=unit.stride, pad=unit.pad, out=dw, accumulate=True)
        rec[tag] = dict(dy=dy, dx=dx, dw=dw, dres=dres, dbeta=sums[0], dgamma=sums[1])
        return dx, dres
    rec = {}
    d_o2, d_idn = unit_bwd(blk.conv3, c3, dout, "conv3", rec)
    d_o1, _ = unit_bwd(blk.conv2, c2, d_o2, "conv2", rec)
    d_x1, _ = unit_bwd(blk.conv1, c1, d_o1, "conv1", rec)
    torch.cuda.synchronize()
    # ---- oracle with hooks on the same tensors ----
    p = {"b." + k: v.requires_grad_(True) for k, v in O.params_from_cuda_module(blk).items()}
    xr = nchw(x).requires_grad_(True)
    keep = {}

    def conv_bn(xx, prefix, stride=1, pad=0, relu=True, residual=None):
        yc = F.conv2d(xx, O.Qf(p[prefix + ".weight"]), stride=stride, padding=pad)
        yc.retain_grad()
        keep[prefix + ".y"] = yc
        y = O.Q(yc)
        y = O.bn_train(y, p[prefix + ".bn.weight"], p[prefix + ".bn.bias"])
        if residual is not None:
            y = y + residual
        z = F.relu(y) if relu else y
        z.retain_grad()
        keep[prefix + ".z"] = z
        return O.Q(z)
    xq = O.Q(xr)
    xq.retain_grad()
    o1 = conv_bn(xq, "b.conv1")
    o2 = conv_bn(o1, "b.conv2", stride=stride, pad=1)
    idn = conv_bn(O.Qb(xq), "b.downsample", stride=stride, relu=False) if ds else xq
    o3 = conv_bn(o2, "b.conv3", residual=idn)
    o3.backward(nchw(dout))
    print("forward out rel %.5f" % rel(out.permute(0, 3, 1, 2), o3))
    for tag in ("conv3", "conv2", "conv1"):
        r = rec[tag]
        gy = keep["b.%s.y" % tag].grad
        print("%s: dy %.5f | dgamma %.5f dbeta %.5f | dw %.5f" % (
            tag, rel(r["dy"].permute(0, 3, 1, 2), gy), rel(r["dgamma"], p["b.%s.bn.weight" % tag].grad),
            rel(r["dbeta"], p["b.%s.bn.bias" % tag].grad), rel(r["dw"].permute(0, 3, 1, 2), p["b.%s.weight" % tag].grad)))
    # the dgrad outputs = the (unrounded-sum) gradient at the producer's output z
    print("dgrad conv3 -> d(o2) %.5f" % rel(rec["conv3"]["dx"].permute(0, 3, 1, 2), O._round_bf16(keep["b.conv2.z"].grad)))
    print("dgrad conv2 -> d(o1) %.5f" % rel(rec["conv2"]["dx"].permute(0, 3, 1, 2), O._round_bf16(keep["b.conv1.z"].grad)))
    # isolate the kernels: feed the ORACLE's (bf16) gradients into the CUDA kernels of one unit
    gz = O._round_bf16(keep["b.conv2.z"].grad)                      # oracle d(o2), bf16-representable
    dz_o = gz.permute(0, 2, 3, 1).contiguous().float().cuda().bfloat16()
    xin, y, z, msss, has_res, _, mask = c2
    dy_o, _, sums_o = K.bn_bwd(y, dz_o, z, msss, blk.conv2.bn.weight, True)
    torch.cuda.synchronize()
    print("conv2 BN backward fed with the oracle's dz: dy %.5f dgamma %.5f dbeta %.5f" % (
        rel(dy_o.permute(0, 3, 1, 2), keep["b.conv2.y"].grad), rel(sums_o[1], p["b.conv2.bn.weight"].grad),
        rel(sums_o[0], p["b.conv2.bn.bias"].grad)))
    gy = keep["b.conv2.y"].grad
    dy_in = gy.permute(0, 2, 3, 1).contiguous().float().cuda().bfloat16()
    dx_o = K.conv2d_dgrad(dy_in, compute_copy(blk.conv2.weight), tuple(xin.shape), stride=stride, pad=1)
    torch.cuda.synchronize()
    print("conv2 dgrad fed with the oracle's dy: d(o1) %.5f" % rel(dx_o.permute(0, 3, 1, 2), O._round_bf16(keep["b.conv1.z"].grad)))


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    main(*a)
