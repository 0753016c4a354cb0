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
    assert AdamW(st, lr=1e-3, weight_decay=0.5).seg_wd.tolist() == [0.5, 0.0, 0.5, 0.0, 0.0]       # 1-d `c` undecayed, f1 / f2 frozen
    lars = LarsMomentumOptimizer(st, lr=0.1, lars_weight_decay=1e-4, exclude_from_weight_decay=("^c$",))
    assert [round(v, 6) for v in lars.seg_wd.tolist()] == [1e-4, 0.0, 1e-4, 0.0, 0.0]
    # frozen tensors have no gradient view: nothing can be accumulated into them by accident
    m = _M()
    ParamStore(m)
    assert m.f1.grad is None and m.f2.grad is None and m.a.grad is not None
