"""The fused flat-buffer optimizer kernels (csrc/optim.cu) against the per-tensor fp64 restatement of the reference's update rules
(oracle/optim.py): several steps over a store with tensors of ragged sizes, per-tensor decay tables, a frozen tensor, grad_scale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w1 = torch.nn.Parameter(torch.randn(96, 33))               # ragged sizes: padding inside the 1024-element slots
        self.b1 = torch.nn.Parameter(0.1 * torch.randn(96))
        self.frozen = torch.nn.Parameter(torch.randn(1, 7, 16), requires_grad=False)
        self.w2 = torch.nn.Parameter(torch.randn(5, 3, 3, 40))          # > 1 slot
        self.g2 = torch.nn.Parameter(1.0 + 0.1 * torch.randn(40))
        self.zero = torch.nn.Parameter(torch.zeros(12, 12))             # ||p|| = 0: LARS falls back to the plain rate


def _setup(seed):
    from passl_b200.core import ParamStore
    torch.manual_seed(seed)
    m = _M().cuda()
    st = ParamStore(m)
    ref = {n: p.detach().double().cpu() for n, p in m.named_parameters()}
    return m, st, ref


def _grads(m, scale):
    out = {}
    for n, p in m.named_parameters():
        if p.requires_grad:
            g = torch.randn_like(p) * scale
            p.grad.copy_(g)
            out[n] = g.double().cpu()
    return out


def _check(m, st, ref, tol):
    for n, p in m.named_parameters():
        got = p.detach().double().cpu()
        err = (got - ref[n]).abs().max().item()
        assert err <= tol * max(1.0, ref[n].abs().max().item()), (n, err)
        assert torch.equal(p.bf16.float(), p.detach().bfloat16().float()), n          # bf16 mirror written by the same kernel
    pad = torch.ones(st.numel, dtype=torch.bool, device="cuda")
    for p, o in zip(st.params, st.offsets):
        pad[o:o + p.numel()] = False
    assert st.master[pad].abs().sum().item() == 0                                    # alignment padding stays zero


@pytest.mark.parametrize("world", [1, 8])
def test_momentum_matches_oracle(world):
    import oracle.optim as O
    from passl_b200.optimizer import Momentum
    m, st, ref = _setup(0)
    opt = Momentum(st, lr=0.03, momentum=0.9, weight_decay=1e-4)
    opt.grad_scale = 1.0 / world                                                      # the mean over ranks folded into the step
    vel = {n: torch.zeros_like(v) for n, v in ref.items()}
    for it in range(4):
        g = _grads(m, 0.5 * world)
        opt.set_lr(0.03 * (1 - 0.1 * it))
        opt.step()
        for n in g:
            ref[n], vel[n] = O.momentum(ref[n], g[n] / world, vel[n], opt.lr, 0.9, 1e-4)
    _check(m, st, ref, 2e-6)


@pytest.mark.parametrize("exclude", [None, ["batch_norm", ".b_0"], ["_0.w_3", "_0.w_1"]])
def test_lars_matches_oracle(exclude):
    import oracle.optim as O
    from passl_b200.optimizer import LarsMomentumOptimizer
    from passl_b200.optimizer.naming import paddle_auto_names
    m, st, ref = _setup(1)
    opt = LarsMomentumOptimizer(st, lr=0.3, momentum=0.9, lars_weight_decay=1e-3, lars_coeff=0.001, exclude_from_weight_decay=exclude)
    auto = dict(zip(st.names, paddle_auto_names(m)))
    wd = {n: (0.0 if any(s in auto[n] for s in (exclude or [])) else 1e-3) for n in ref}
    vel = {n: torch.zeros_like(v) for n, v in ref.items()}
    for it in range(4):
        g = _grads(m, 0.2)
        opt.step()
        for n in g:
            ref[n], vel[n] = O.lars_momentum(ref[n], g[n], vel[n], 0.3, 0.9, wd[n], 0.001, 0.0)
    _check(m, st, ref, 5e-6)
    assert ref["zero"].abs().max() > 0                                                # the zero-norm tensor moved with the plain rate


def test_adamw_matches_oracle():
    import oracle.optim as O
    from passl_b200.optimizer import AdamW
    m, st, ref = _setup(2)
    ratio = {"w1": 0.5, "b1": 0.5}
    opt = AdamW(st, lr=1.5e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.05, no_decay=["^zero$"], one_dim_no_decay=True,
                lr_ratio=lambda n, p: ratio.get(n, 1.0))
    wd = {n: (0.0 if (v.dim() <= 1 or n == "zero") else 0.05) for n, v in ref.items()}
    mom = {n: (torch.zeros_like(v), torch.zeros_like(v)) for n, v in ref.items()}
    for it in range(5):
        g = _grads(m, 1.0)
        opt.step()
        for n in g:
            ref[n], m1, m2 = O.adamw(ref[n], g[n], mom[n][0], mom[n][1], 1.5e-3, 0.9, 0.95, 1e-8, wd[n], it + 1, ratio.get(n, 1.0))
            mom[n] = (m1, m2)
    _check(m, st, ref, 5e-6)
    assert np.isfinite(st.master.sum().item())
