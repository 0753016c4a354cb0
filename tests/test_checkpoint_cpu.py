"""Paddle-checkpoint name / layout mapping (passl_b200/utils/checkpoint.py) checked against the reference classes themselves: the
reference's ResNet-50 (resnetimagenet.py) and NonLinearNeckV1 (base_neck.py) are constructed over the paddle shim — no weights
needed, only their parameter / buffer names and shapes — and must equal what `moco_to_paddle` emits."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _model():
    from passl_b200.modeling import build_model
    from passl_b200.utils.config import get_config
    cfg = get_config(os.path.join(os.path.dirname(HERE), "configs/moco/moco_v2_r50.yaml"), ["model.K=512"])
    return build_model(dict(cfg.model))


def test_roundtrip_and_file_container(tmp_path):
    from passl_b200.utils import checkpoint as C
    torch.manual_seed(0)
    a, b = _model(), _model()
    for p in a.parameters():
        torch.nn.init.normal_(p, std=0.1)
    for n, buf in a.named_buffers():
        if buf.dtype.is_floating_point:
            buf.normal_()
    a.queue_ptr.fill_(128)
    state = C.moco_to_paddle(a)
    assert state["queue"].shape == (128, 512) and state["encoder_q.0.conv1.weight"].shape == (64, 3, 7, 7)
    assert state["encoder_q.1.mlp.0.weight"].shape == (2048, 2048) and state["encoder_q.1.mlp.2.weight"].shape == (2048, 128)
    path = str(tmp_path / "epoch_1.pdparams")
    C.save_pdparams(state, path)
    C.moco_from_paddle(b, C.load_pdparams(path))
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("stem.weight"):
            assert torch.equal(sa[k][:, :147], sb[k][:, :147]) and sb[k][:, 147:].abs().sum() == 0
        else:
            assert torch.equal(sa[k], sb[k]), k


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (build container only)")
def test_names_and_shapes_equal_the_reference_classes():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    import make_golden_models as M
    import paddle_shim  # noqa: F401
    make_golden.setup()
    M.extend_shim()
    nn = sys.modules["paddle.nn"]

    class Conv2D(nn.Layer):
        def __init__(self, i, o, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias_attr=None, **kw):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.zeros(o, i, kernel_size, kernel_size))

    class BatchNorm2D(nn.Layer):
        def __init__(self, c, **kw):
            super().__init__()
            self.weight, self.bias = torch.nn.Parameter(torch.ones(c)), torch.nn.Parameter(torch.zeros(c))
            self.register_buffer("_mean", torch.zeros(c))
            self.register_buffer("_variance", torch.ones(c))

    class MaxPool2D(nn.Layer):
        def __init__(self, *a, **k):
            super().__init__()
    nn.Conv2D, nn.BatchNorm2D, nn.MaxPool2D = Conv2D, BatchNorm2D, MaxPool2D
    rn = importlib.import_module("passl_v110.modeling.backbones.resnetimagenet")
    necks = importlib.import_module("passl_v110.modeling.necks.base_neck")
    ref_backbone = rn.ResNet(rn.BottleneckBlock, 50, num_classes=0, with_pool=False)
    ref_neck = necks.NonLinearNeckV1(in_channels=2048, hid_channels=2048, out_channels=128)
    want = {}
    for pre, mod in (("encoder_q.0.", ref_backbone), ("encoder_q.1.", ref_neck)):
        for k, v in mod.state_dict().items():
            want[pre + k] = tuple(v.shape)
    from passl_b200.utils import checkpoint as C
    got = {k: tuple(v.shape) for k, v in C.moco_to_paddle(_model()).items() if k.startswith("encoder_q.")}
    assert set(got) == set(want), (sorted(set(want) - set(got))[:5], sorted(set(got) - set(want))[:5])
    for k in want:
        assert got[k] == want[k], (k, got[k], want[k])
