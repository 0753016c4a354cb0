"""The C-ABI library loads and exports every symbol include/passl_b200.h declares (no compute without a GPU), and the
host-side mirrors of the reference interface behave like the reference (registry errors, contract violations)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from passl_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "passl_b200.h")).read()
    declared = set(re.findall(r"\b(passl_b200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.passl_b200_version() >= 100


def test_kernels_refuse_cpu_tensors():
    import torch
    from passl_b200 import kernels as K, _lib
    with pytest.raises(_lib.PasslB200Error):
        K.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    with pytest.raises(_lib.PasslB200Error):
        K.l2norm_fwd(torch.zeros(4, 128))


def test_registry_contract():
    from passl_b200.modeling import MODELS, BACKBONES, NECKS, HEADS, build_backbone
    from passl_b200.utils.registry import Registry, build_from_config
    for name in ["MoCo"]:
        assert name in MODELS
    for name in ["ResNet", "ResNetsimclr"]:
        assert name in BACKBONES
    for name in ["LinearNeck", "NonLinearNeckV1", "NonLinearNeckfc3"]:
        assert name in NECKS
    assert "ContrastiveHead" in HEADS
    with pytest.raises(KeyError):
        build_backbone(dict(name="NoSuchNet"))
    with pytest.raises(TypeError):
        build_from_config(["not", "a", "dict"], BACKBONES)
    with pytest.raises(KeyError):
        build_from_config(dict(depth=50), BACKBONES)
    r = Registry("X")
    r.register(name="a")(int)
    with pytest.raises(AssertionError):
        r.register(name="a")(float)


def test_resnet50_parameter_count_matches_reference_topology():
    from passl_b200.modeling import build_backbone, build_neck
    net = build_backbone(dict(name="ResNet", depth=50))
    n = sum(p.numel() for p in net.parameters()) - 64 * 5        # stem K padding 147 -> 152
    assert n == 23508032                                           # torchvision / paddle.vision resnet50 without fc
    neck = build_neck(dict(name="NonLinearNeckV1", in_channels=2048, hid_channels=2048, out_channels=128))
    assert sum(p.numel() for p in neck.parameters()) == 2048 * 2048 + 2048 + 2048 * 128 + 128
