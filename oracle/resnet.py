"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the ResNet-50 backbone + necks: a torch-CPU restatement of
passl_v110/modeling/backbones/resnetimagenet.py:93-246 (BottleneckBlock :93-148, ResNet :150-246) and
necks/base_neck.py:43-94,209-237 in NCHW float32/float64, driven by explicit parameter dicts so the same weights can be
loaded into the CUDA path.  Paddle semantics restated: BatchNorm train mode normalises with the biased batch variance,
eps 1e-5; running = 0.9*running + 0.1*batch (biased variance, as paddle's CPU batch_norm kernel); use_global_stats=True
normalises with the running statistics (passl_v110/modules/freeze.py:17-23).
Parity status: pinned (tests/test_oracle_resnet_cpu.py, tests/test_oracle_necks_cpu.py) against (1) golden outputs of the
reference's own ResNet / neck classes executed over the torch-backed paddle shim (reduced-width network built by the reference's
`_make_layer` + `BottleneckBlock`; NonLinearNeckV1 / LinearNeck / NonLinearNeckfc3), and (2) torchvision's resnet50 (same v1.5
topology) with copied weights, train-mode and eval-mode BatchNorm, float64, 1e-9.  Paddle itself cannot run here.
"""
import torch
import torch.nn.functional as F


# ---- quantisation-matched mode (q=True) -----------------------------------------------------------------------------------
# The CUDA path keeps activations and activation gradients in bf16 between fused units (fp32 accumulation inside a unit).  With
# q=True this oracle rounds to bf16 at exactly those points, forward AND backward, so that what is left between the two is
# accumulation order — the comparison then isolates kernel defects from quantisation noise (which BatchNorm over small batches
# amplifies through 50 layers).  Rounding points of the CUDA path (passl_b200/nn/layers.py, modeling/backbones/resnet.py):
#   forward : input pixels; conv output y (statistics are taken from the rounded values); unit output z = relu(bn(y) [+ res]);
#             max-pool / avg-pool output; fc outputs (bf16 ones)
#   backward: dz arriving at a unit (sum of its consumers' contributions, rounded once); dy leaving the BN backward;
#             the downsample branch's dx before the conv1 dgrad accumulates onto it
def _round_bf16(x):
    return x.float().bfloat16().to(x.dtype)


class _RoundBoth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _round_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _round_bf16(g)


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _round_bf16(g)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _round_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _ReluGivenMask(torch.autograd.Function):
    """relu(x) forward; the backward uses a GIVEN 0/1 mask instead of (x > 0).  A ReLU gradient is discontinuous: where the
    pre-activation is within rounding noise of zero the two implementations can take different sides, and every such element
    changes its gradient by 100 % — a forward difference of relative size e flips ~0.4 e of the masks and moves the backward
    by ~sqrt(0.4 e) (1e-3 forward -> 2 % backward).  Parity tests of a backward pass therefore hand the CUDA path's masks to the
    oracle; the forward comparison (and the masks themselves, through it) stays independent."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return torch.relu(x)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask, None


def Q(x, on=True):
    return _RoundBoth.apply(x) if on else x


def Qf(x, on=True):
    """forward-only rounding: the bf16 mirror of an fp32 master weight (its gradient is accumulated in fp32), or an fp32
    embedding cast to bf16 for a tensor-core kernel whose input gradient comes back in fp32"""
    return _RoundFwd.apply(x) if on else x


def Qb(x, on=True):
    return _RoundBwd.apply(x) if on else x


def bn_train(x, gamma, beta, eps=1e-5, stats=None, use_global_stats=False, running=None):
    dims = [0] + list(range(2, x.dim()))
    if use_global_stats:
        mean, var = running
    else:
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if stats is not None:
            stats.append((mean.detach().clone(), var.detach().clone()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    y = (x - mean.reshape(shape)) / torch.sqrt(var.reshape(shape) + eps)
    if gamma is not None:
        y = y * gamma.reshape(shape) + beta.reshape(shape)
    return y


def conv_bn(x, p, prefix, stride=1, pad=0, relu=True, residual=None, use_global_stats=False, q=False, relu_mask=None):
    """p[prefix+'.weight'] is [Cout, Cin, R, S] (NCHW convention)."""
    y = Q(F.conv2d(x, Qf(p[prefix + ".weight"], q), stride=stride, padding=pad), q)
    running = (p.get(prefix + ".bn._mean"), p.get(prefix + ".bn._variance"))
    y = bn_train(y, p[prefix + ".bn.weight"], p[prefix + ".bn.bias"], use_global_stats=use_global_stats, running=running)
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y) if relu_mask is None else _ReluGivenMask.apply(y, relu_mask.to(y.dtype))
    return Q(y, q)


def bottleneck(x, p, prefix, stride, has_ds, ugs=False, q=False, masks=(None, None, None)):
    """resnetimagenet.py:133-148.  masks: optional ReLU masks (NCHW 0/1) of the three units for the backward (_ReluGivenMask)."""
    out = conv_bn(x, p, prefix + ".conv1", use_global_stats=ugs, q=q, relu_mask=masks[0])
    out = conv_bn(out, p, prefix + ".conv2", stride=stride, pad=1, use_global_stats=ugs, q=q, relu_mask=masks[1])
    identity = conv_bn(Qb(x, q), p, prefix + ".downsample", stride=stride, relu=False, use_global_stats=ugs, q=q) if has_ds else x
    return conv_bn(out, p, prefix + ".conv3", relu=True, residual=identity, use_global_stats=ugs, q=q, relu_mask=masks[2])


def resnet_forward(img, p, layers=(3, 4, 6, 3), stem_maxpool=True, with_pool=False, ugs=False, prefix="", q=False, masks=None):
    """img NCHW; returns NCHW feature map (or [B, C] when with_pool).  resnetimagenet.py:232-246.
    masks: optional dict(stem=mask, blocks=[(m1, m2, m3), ...]) of ReLU masks for the backward (_ReluGivenMask)."""
    if q:
        img = _round_bf16(img)
    x = conv_bn(img, p, prefix + "stem", stride=2, pad=3, use_global_stats=ugs, q=q, relu_mask=None if masks is None else masks["stem"])
    if stem_maxpool:
        x = Q(F.max_pool2d(x, kernel_size=3, stride=2, padding=1), q)      # forward no-op (max commutes with rounding)
    inplanes, bi = 64, 0
    for i, (planes, n) in enumerate(zip([64, 128, 256, 512], layers)):
        for b in range(n):
            s = (1 if i == 0 else 2) if b == 0 else 1
            has_ds = b == 0 and (s != 1 or inplanes != planes * 4)
            x = bottleneck(x, p, prefix + "blocks.%d" % bi, s, has_ds, ugs, q=q,
                           masks=(None, None, None) if masks is None else masks["blocks"][bi])
            inplanes = planes * 4
            bi += 1
    if with_pool:
        x = Q(x.mean(dim=(2, 3)), q)
    return x


def _neck_in(feat, with_avg_pool, q):
    if with_avg_pool and feat.dim() == 4:
        return Q(feat.mean(dim=(2, 3)), q)
    return feat.reshape(feat.shape[0], -1)


def neck_linear(feat, p, prefix="", with_avg_pool=True, q=False):
    """LinearNeck (base_neck.py:43-64): avgpool -> fc (fp32 output; its gradient is cast to bf16 for the dgrad / wgrad GEMMs)."""
    x = _neck_in(feat, with_avg_pool, q)
    return Qb(F.linear(x, Qf(p[prefix + "fc.weight"], q), p[prefix + "fc.bias"]), q)


def neck_v1(feat, p, prefix="", with_avg_pool=True, q=False):
    """NonLinearNeckV1 (base_neck.py:67-94): avgpool -> fc -> relu -> fc; weights here are [out, in]."""
    x = _neck_in(feat, with_avg_pool, q)
    x = Q(F.relu(F.linear(x, Qf(p[prefix + "fc1.weight"], q), p[prefix + "fc1.bias"])), q)
    return Qb(F.linear(x, Qf(p[prefix + "fc2.weight"], q), p[prefix + "fc2.bias"]), q)


def neck_fc3(feat, p, prefix="", with_avg_pool=False, q=False):
    """NonLinearNeckfc3 (base_neck.py:209-237) incl. the trailing l2_normalize(hidden, -1)."""
    x = _neck_in(feat, with_avg_pool, q)
    for i in (1, 2, 3):
        x = Q(F.linear(x, Qf(p[prefix + "fc%d.weight" % i], q), p[prefix + "fc%d.bias" % i]), q)
        x = bn_train(x, p[prefix + "bn%d.bn.weight" % i], p[prefix + "bn%d.bn.bias" % i])
        x = Q(F.relu(x), q) if i < 3 else Qb(x, q)        # the last BN writes fp32; its gradient arrives in bf16
    return x / torch.sqrt((x * x).sum(-1, keepdim=True) + 1e-12)


def params_from_cuda_module(module, dtype=torch.float64, bf16_round=True):
    """Export a passl_b200 ResNet / neck module's parameters into the NCHW dict this oracle consumes.
    bf16_round: round weights to bf16 first (what the tensor cores actually multiply)."""
    out = {}
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        v = t.detach().float().cpu()
        if name.endswith("stem.weight"):
            v = v[:, :147].reshape(64, 7, 7, 3)
        if v.dim() == 4:                                    # [Cout, R, S, Cin] -> [Cout, Cin, R, S]
            if bf16_round:
                v = v.bfloat16().float()
            v = v.permute(0, 3, 1, 2).contiguous()
        elif v.dim() == 2 and bf16_round:
            v = v.bfloat16().float()
        out[name] = v.to(dtype)
    return out
