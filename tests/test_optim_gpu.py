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


def test_grad_norm_finite_clip_and_skip():
    """One-pass global norm + finite check (grad_clip.py:30-84, grad_scaler.py:48-87): control word on the device, consumed by the
    fused update kernels — clip coefficient, unscale, and the skipped step on a non-finite gradient."""
    import torch.nn as nn
    from passl_b200.core import ParamStore
    from passl_b200.optimizer import GradControl, Momentum
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(300, 200), nn.Linear(200, 10)).cuda()
    store = ParamStore(net)
    opt = Momentum(store, lr=0.1, momentum=0.0, weight_decay=0.0)
    store.grad.normal_()
    store.grad.mul_((store.master != 0).float())          # alignment padding carries no gradient
    g = store.grad.clone()
    w0 = store.master.clone()
    norm = g.norm().item()
    # 1. clip active: update = lr * g * clip_norm / (norm + 1e-6)
    opt.set_grad_clip(dict(name="ClipGradByGlobalNorm", clip_norm=1.0))
    opt.step()
    torch.cuda.synchronize()
    assert abs(opt.grad_control.global_norm.item() - norm) < 1e-4 * norm
    exp = w0 - 0.1 * g * (1.0 / (norm + 1e-6))
    assert torch.allclose(store.master, exp, rtol=1e-5, atol=1e-7)
    # 2. clip inactive (norm below the threshold): plain step; clip_norm_max caps the coefficient when always_clip
    store.master.copy_(w0); store.refresh_bf16()
    opt.set_grad_clip(dict(clip_norm=10.0 * norm))
    opt.step()
    assert torch.allclose(store.master, w0 - 0.1 * g, rtol=1e-5, atol=1e-7)
    store.master.copy_(w0)
    opt.set_grad_clip(dict(clip_norm=10.0 * norm, always_clip=True, clip_norm_max=2.0))
    opt.step()
    assert torch.allclose(store.master, w0 - 0.1 * g * 2.0, rtol=1e-5, atol=1e-7)
    # 3. loss-scaled gradients: unscale folded into the multiplier
    store.master.copy_(w0)
    opt.grad_control = GradControl(store, loss_scale=1024.0)
    store.grad.copy_(g * 1024.0)
    opt.step()
    assert torch.allclose(store.master, w0 - 0.1 * g, rtol=1e-5, atol=1e-7)
    # 4. a non-finite gradient: the whole step is skipped on the device, found_inf raised
    store.master.copy_(w0)
    store.grad.copy_(g)
    store.grad[5] = float("inf")
    opt.step()
    torch.cuda.synchronize()
    assert opt.grad_control.found_inf.item() == 1.0
    assert torch.equal(store.master, w0)
    store.grad[5] = float("nan")
    opt.step()
    assert opt.grad_control.found_inf.item() == 1.0 and torch.equal(store.master, w0)
    # the scratch word resets itself: a clean gradient afterwards steps normally
    store.grad.copy_(g * 1024.0)
    opt.step()
    assert opt.grad_control.found_inf.item() == 0.0
    assert torch.allclose(store.master, w0 - 0.1 * g, rtol=1e-5, atol=1e-7)


def test_grad_scaler_surface():
    from passl_b200.core.grad_scaler import GradScaler
    s = GradScaler(enable=False)
    x = torch.ones(1, device="cuda", requires_grad=True)
    assert s.scale(x) is x
    s2 = GradScaler(enable=True, init_loss_scaling=8.0)
    y = s2.scale(x.sum())
    y.backward()
    assert x.grad.item() == 8.0


def test_shadow_ema_matches_the_reference_rule():
    """passl/models/utils/ema.py:18-97: decay = min(decay, (1 + t) / (10 + t)) when thres_steps; apply_shadow / restore swap."""
    import torch.nn as nn
    from passl_b200.core import ParamStore
    from passl_b200.models.utils_ema import EMA
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(64, 64), nn.Linear(64, 8)).cuda()
    store = ParamStore(net)
    ema = EMA(store, decay=0.999)
    ema.register()
    ref = store.master.clone().double()
    for t in range(4):
        store.master.add_(0.01 * torch.randn_like(store.master))
        d = ema.update()
        assert d == min(0.999, (1 + t) / (10 + t))
        ref = d * ref + (1 - d) * store.master.double()
    assert torch.allclose(ema.state_dict()["shadow"].double(), ref, rtol=1e-6, atol=1e-7)
    live = store.master.clone()
    ema.apply_shadow()
    assert torch.allclose(store.master.double(), ref, rtol=1e-6, atol=1e-7) and torch.equal(store.bf16, store.master.bfloat16())
    ema.restore()
    assert torch.equal(store.master, live)
