"""Cross-oracle: the ResNet-50 restatement in oracle/resnet.py against torchvision's resnet50 — the same
v1.5 topology as paddle.vision / resnetimagenet.py (stride on the 3x3 conv, 1x1 stride-s downsample) written independently.
Train-mode BatchNorm (batch statistics), float64, copied weights: features must agree to rounding error."""
import pytest
import torch

tv = pytest.importorskip("torchvision")


def _export_torchvision(net):
    """torchvision names -> the parameter dict of oracle.resnet (passl_b200 module names)."""
    p = {}
    sd = {k: v.double() for k, v in net.state_dict().items()}

    def unit(dst, conv, bn):
        p[dst + ".weight"] = sd[conv + ".weight"]
        p[dst + ".bn.weight"], p[dst + ".bn.bias"] = sd[bn + ".weight"], sd[bn + ".bias"]
        p[dst + ".bn._mean"], p[dst + ".bn._variance"] = sd[bn + ".running_mean"], sd[bn + ".running_var"]
    unit("stem", "conv1", "bn1")
    bi = 0
    for li, n in enumerate([3, 4, 6, 3], start=1):
        for b in range(n):
            src = "layer%d.%d" % (li, b)
            for k in (1, 2, 3):
                unit("blocks.%d.conv%d" % (bi, k), "%s.conv%d" % (src, k), "%s.bn%d" % (src, k))
            if b == 0:
                unit("blocks.%d.downsample" % bi, src + ".downsample.0", src + ".downsample.1")
            bi += 1
    return p


def test_oracle_resnet50_matches_torchvision_train_mode():
    import oracle.resnet as OR
    torch.manual_seed(0)
    net = tv.models.resnet50(weights=None).double()
    for m in net.modules():                                  # non-trivial affine parameters / running statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            torch.nn.init.uniform_(m.weight, 0.5, 1.5)
            torch.nn.init.normal_(m.bias, std=0.1)
    net.train()
    img = torch.randn(4, 3, 64, 64, dtype=torch.float64)
    p = _export_torchvision(net)
    x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
    for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
        x = layer(x)
    ref_map = x
    ref_pool = torch.flatten(net.avgpool(x), 1)
    got_map = OR.resnet_forward(img, p)
    got_pool = OR.resnet_forward(img, p, with_pool=True)
    torch.testing.assert_close(got_map, ref_map, rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(got_pool, ref_pool, rtol=1e-9, atol=1e-9)


def test_oracle_resnet50_global_stats_matches_torchvision_eval_mode():
    """use_global_stats (freeze.py:17-23, MoCo key encoder) == eval-mode BatchNorm."""
    import oracle.resnet as OR
    torch.manual_seed(1)
    net = tv.models.resnet50(weights=None).double()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            torch.nn.init.uniform_(m.running_var, 0.5, 1.5)
            torch.nn.init.normal_(m.running_mean, std=0.2)
    net.eval()
    img = torch.randn(2, 3, 64, 64, dtype=torch.float64)
    with torch.no_grad():
        x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            x = layer(x)
        got = OR.resnet_forward(img, _export_torchvision(net), ugs=True)
    torch.testing.assert_close(got, x, rtol=1e-9, atol=1e-9)


def test_oracle_matches_reference_resnet_code_reduced_width():
    """oracle/resnet.py against golden output of the reference's OWN ResNet code (resnetimagenet.py `_make_layer` +
    `BottleneckBlock` + stem, reduced widths; tests/golden/make_golden_models.py gen_resnet_layer)."""
    import os
    import numpy as np
    import oracle.resnet as OR
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_resnet_layers.npz"))
    p = {}
    blocks = {"layer1.0": 0, "layer1.1": 1, "layer2.0": 2, "layer2.1": 3}
    for key in g.files:
        if not key.startswith("w_"):
            continue
        name, v = key[2:], torch.from_numpy(g[key])
        if name == "conv1.weight":
            p["stem.weight"] = v
        elif name.startswith("bn1."):
            p["stem.bn." + name[4:]] = v
        else:
            lay, rest = name[:8], name[9:]
            pre = "blocks.%d." % blocks[lay]
            if rest.startswith("downsample.0."):
                p[pre + "downsample.weight"] = v
            elif rest.startswith("downsample.1."):
                p[pre + "downsample.bn." + rest[13:]] = v
            elif rest.startswith("conv"):
                p[pre + rest] = v                                   # conv{k}.weight
            else:                                                   # bn{k}.weight / bias
                p[pre + "conv" + rest[2] + ".bn." + rest[4:]] = v
    y = OR.resnet_forward(torch.from_numpy(g["x"]), p, layers=(2, 2))
    np.testing.assert_allclose(y.numpy(), g["y"], rtol=1e-9, atol=1e-11)
