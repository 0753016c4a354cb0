"""Host-side logic of the flat-buffer optimizers that needs no GPU: which tensors take a step.  The reference's group builder drops
stop_gradient parameters (passl/optimizer/__init__.py:88-91,117), so a frozen tensor sees neither update nor weight decay."""
import torch

from passl_b200.core import ParamStore
from passl_b200.optimizer import AdamW, LarsMomentumOptimizer, Momentum, trainable_ranges


class _M(torch.nn.Module):
    def __init__(self, frozen=("f1", "f2")):
        super().__init__()
        self.a = torch.nn.Parameter(torch.randn(40, 50))            # 2000 -> 2048 slots
        self.f1 = torch.nn.Parameter(torch.randn(3, 700))           # 2100 -> 3072
        self.b = torch.nn.Parameter(torch.randn(64, 33))            # 2112 -> 3072
        self.c = torch.nn.Parameter(torch.randn(7))                 # 7 -> 1024
        self.f2 = torch.nn.Parameter(torch.randn(1, 5, 8))          # 40 -> 1024
        for n in frozen:
            getattr(self, n).requires_grad = False


def test_trainable_ranges_skip_frozen_runs():
    st = ParamStore(_M())
    assert st.offsets == [0, 2048, 5120, 8192, 9216] and st.numel == 10240
    assert trainable_ranges(st) == [(0, 2048), (5120, 4096)]
    assert trainable_ranges(ParamStore(_M(frozen=()))) == [(0, 10240)]                  # nothing frozen: one launch over everything
    assert trainable_ranges(ParamStore(_M(frozen=("a",)))) == [(2048, 8192)]
    opt = Momentum(st, lr=0.1, weight_decay=1e-4)
    assert opt._ranges == [(0, 2048), (5120, 4096)] and all(o % 4 == 0 and n % 4 == 0 for o, n in opt._ranges)


def test_weight_decay_tables_zero_frozen_tensors():
    st = ParamStore(_M())
    # MAE rule: 1-d `c` undecayed; default (MoCo v3 / CLIP recipes): every trainable tensor decays; f1 / f2 frozen either way
    assert AdamW(st, lr=1e-3, weight_decay=0.5, one_dim_no_decay=True).seg_wd.tolist() == [0.5, 0.0, 0.5, 0.0, 0.0]
    assert AdamW(st, lr=1e-3, weight_decay=0.5).seg_wd.tolist() == [0.5, 0.0, 0.5, 0.5, 0.0]
    assert AdamW(st, lr=1e-3, weight_decay=0.5, no_weight_decay_name=["b"]).seg_wd.tolist() == [0.5, 0.0, 0.0, 0.5, 0.0]
    o = AdamW(st, betas="(0.9, 0.95)", eps=1e-6, use_master_param=True, exp_avg_force_fp32=True)
    assert (o.beta1, o.beta2, o.eps) == (0.9, 0.95, 1e-6)
    lars = LarsMomentumOptimizer(st, lr=0.1, lars_weight_decay=1e-4, exclude_structured=("^c$",))
    assert [round(v, 6) for v in lars.seg_wd.tolist()] == [1e-4, 0.0, 1e-4, 0.0, 0.0]
    lars = LarsMomentumOptimizer(st, lr=0.1, lars_weight_decay=1e-4)                 # fluid default: nothing excluded but frozen
    assert [round(v, 6) for v in lars.seg_wd.tolist()] == [1e-4, 0.0, 1e-4, 1e-4, 0.0]
    # frozen tensors have no gradient view: nothing can be accumulated into them by accident
    m = _M()
    ParamStore(m)
    assert m.f1.grad is None and m.f2.grad is None and m.a.grad is not None


def test_lars_exclude_list_matches_paddle_generated_names():
    """`exclude_from_weight_decay` is a list of substrings of Paddle's generated parameter names (optimizer/naming.py).  On the SimCLR
    encoder (ResNet-50 + fc3 neck) the SimCLR YAML's list (configs/simclr/simclr_r50_IM.yaml:120) therefore excludes nothing, while
    the BYOL YAML's list (configs/moco_byol/moco_byol_r50_IM.yaml:122) excludes every BatchNorm tensor and every bias."""
    from collections import Counter
    from passl_b200.modeling import build_model
    from passl_b200.optimizer.naming import paddle_auto_names
    model = build_model(dict(name="SimCLR", backbone=dict(name="ResNet", depth=50, with_pool=True),
                             neck=dict(name="NonLinearNeckfc3", in_channels=2048, hid_channels=2048, out_channels=128, with_avg_pool=False),
                             head=dict(name="SimCLRContrastiveHead", temperature=0.1)))
    names = paddle_auto_names(model.encoder)
    kinds = Counter((n.rsplit("_", 2)[0], n.rsplit(".", 1)[1]) for n in names)
    assert kinds == {("conv2d", "w_0"): 53, ("batch_norm2d", "w_0"): 53, ("batch_norm2d", "b_0"): 53, ("linear", "w_0"): 3,
                     ("linear", "b_0"): 3, ("batch_norm1d", "w_0"): 3, ("batch_norm1d", "b_0"): 3}
    assert names[0] == "conv2d_0.w_0" and len(set(names)) == len(names) == 171
    st = ParamStore(model.encoder)
    simclr = LarsMomentumOptimizer(st, exclude_from_weight_decay=["scale", "offset", ".bias"])
    byol = LarsMomentumOptimizer(st, exclude_from_weight_decay=["batch_norm", ".b_0"])
    assert int((simclr.seg_wd == 0).sum()) == 0
    assert int((byol.seg_wd == 0).sum()) == 2 * 53 + 2 * 3 + 3
    by_name = dict(zip(st.names, byol.seg_wd.tolist()))
    assert by_name["0.blocks.0.conv1.weight"] > 0 and by_name["0.blocks.0.conv1.bn.weight"] == 0 and by_name["1.fc1.bias"] == 0
