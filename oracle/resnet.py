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


def conv_bn(x, p, prefix, stride=1, pad=0, relu=True, residual=None, use_global_stats=False):
    """p[prefix+'.weight'] is [Cout, Cin, R, S] (NCHW convention)."""
    y = F.conv2d(x, p[prefix + ".weight"], stride=stride, padding=pad)
    running = (p.get(prefix + ".bn._mean"), p.get(prefix + ".bn._variance"))
    y = bn_train(y, p[prefix + ".bn.weight"], p[prefix + ".bn.bias"], use_global_stats=use_global_stats, running=running)
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def bottleneck(x, p, prefix, stride, has_ds, ugs=False):
    """resnetimagenet.py:133-148"""
    out = conv_bn(x, p, prefix + ".conv1", use_global_stats=ugs)
    out = conv_bn(out, p, prefix + ".conv2", stride=stride, pad=1, use_global_stats=ugs)
    identity = conv_bn(x, p, prefix + ".downsample", stride=stride, relu=False, use_global_stats=ugs) if has_ds else x
    return conv_bn(out, p, prefix + ".conv3", relu=True, residual=identity, use_global_stats=ugs)


def resnet_forward(img, p, layers=(3, 4, 6, 3), stem_maxpool=True, with_pool=False, ugs=False, prefix=""):
    """img NCHW; returns NCHW feature map (or [B, C] when with_pool).  resnetimagenet.py:232-246."""
    x = conv_bn(img, p, prefix + "stem", stride=2, pad=3, use_global_stats=ugs)
    if stem_maxpool:
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    inplanes, bi = 64, 0
    for i, (planes, n) in enumerate(zip([64, 128, 256, 512], layers)):
        for b in range(n):
            s = (1 if i == 0 else 2) if b == 0 else 1
            has_ds = b == 0 and (s != 1 or inplanes != planes * 4)
            x = bottleneck(x, p, prefix + "blocks.%d" % bi, s, has_ds, ugs)
            inplanes = planes * 4
            bi += 1
    if with_pool:
        x = x.mean(dim=(2, 3))
    return x


def neck_v1(feat, p, prefix="", with_avg_pool=True):
    """NonLinearNeckV1 (base_neck.py:67-94): avgpool -> fc -> relu -> fc; weights here are [out, in]."""
    x = feat.mean(dim=(2, 3)) if (with_avg_pool and feat.dim() == 4) else feat.reshape(feat.shape[0], -1)
    x = F.relu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(x, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def neck_fc3(feat, p, prefix=""):
    """NonLinearNeckfc3 (base_neck.py:209-237) incl. the trailing l2_normalize(hidden, -1)."""
    x = feat.reshape(feat.shape[0], -1)
    for i in (1, 2, 3):
        x = F.linear(x, p[prefix + "fc%d.weight" % i], p[prefix + "fc%d.bias" % i])
        x = bn_train(x, p[prefix + "bn%d.bn.weight" % i], p[prefix + "bn%d.bn.bias" % i])
        if i < 3:
            x = F.relu(x)
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
