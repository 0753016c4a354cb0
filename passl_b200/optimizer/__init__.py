"""Optimizers over ParamStore flat buffers — one fused kernel per step.

API follows the reference's torch-style optimizers (passl/optimizer/optimizer.py:32-233): ``step()``, ``clear_grad()``,
``lr`` attribute driven by an LR scheduler (optimizer/lr.py), ``state_dict()``.  Weight-decay exclusion is a per-tensor table:
AdamW follows the reference's param-group rule (1-d tensors and `no_decay` names undecayed, frozen tensors skipped —
passl/optimizer/__init__.py:124-215, tasks/ssl/mae/util/optim_factory.py:21-38); LARS follows v110's
``exclude_from_weight_decay`` substring list over Paddle's generated parameter names (optimizer/naming.py).
"""
import re

import torch

from .. import _lib, kernels as K
from ..distributed import get_world_size
from .lr import (CosineAnnealingDecay, LinearWarmup, MultiStepDecay, build_lr_scheduler,  # noqa: F401
                 build_lr_scheduler_simclr, build_lr_scheduler_v2)


def _wd_table(store, weight_decay, exclude=None, exclude_structured=None):
    """Per-tensor decay table.  `exclude`: substrings of the Paddle-style auto name (`conv2d_3.w_0`, `batch_norm2d_3.b_0`,
    `linear_0.b_0`), the reference's LARS semantics — see optimizer/naming.py; `exclude_structured`: regexes on the state_dict
    key (`blocks.3.conv1.bn.weight`), for callers who want to name tensors the PyTorch way."""
    from .naming import paddle_auto_names
    subs = list(exclude or [])
    pats = [re.compile(p) for p in (exclude_structured or [])]
    auto = dict(zip(store.names, paddle_auto_names(store.module))) if subs else {}

    def fn(name, p):
        if not p.requires_grad:                      # frozen tensors take no step at all (passl/optimizer/__init__.py:88-91,117)
            return 0.0
        if any(s in auto[name] for s in subs) or any(r.search(name) for r in pats):
            return 0.0
        return weight_decay
    return store.segment_values(fn)


def trainable_ranges(store):
    """[(offset, numel)] of the maximal runs of trainable tensors in the flat buffer (alignment padding included).  Frozen
    tensors (pos-embeddings, stop_grad_conv1, frozen stages) live in the same buffer but must see neither the gradient step nor
    weight decay — the reference's group builder drops stop_gradient parameters (passl/optimizer/__init__.py:88-91,117) and
    adamw.py:57-59 skips parameters without a gradient."""
    runs, start = [], None
    bounds = list(store.offsets[1:]) + [store.numel]
    for p, o, e in zip(store.params, store.offsets, bounds):
        if p.requires_grad:
            start = o if start is None else start
        elif start is not None:
            runs.append((start, o - start))
            start = None
    if start is not None:
        runs.append((start, store.numel - start))
    return runs


class GradControl:
    """ClipGradByGlobalNorm (passl/core/grad_clip.py:30-84) and the scaler's check_finite_and_unscale (passl/core/grad_scaler.py:48-87)
    as ONE read pass over the flat gradient buffer (passl_b200_grad_norm_finite): the result is a 3-float control word on the
    device, {multiplier, found_inf, global_norm}, that the fused optimizer kernels consume — no host round trip, no second pass
    over the gradients.  `no_clip_list` / per-parameter `need_clip=False` of the reference are not built (no headline recipe uses
    them): every tensor of the store takes part in the norm."""

    def __init__(self, store, clip_norm=None, clip_norm_max=None, always_clip=False, loss_scale=1.0):
        self.store = store
        self.clip_norm = float(clip_norm) if clip_norm else 0.0
        self.clip_norm_max = float(clip_norm_max) if clip_norm_max else 0.0
        self.always_clip = bool(always_clip)
        self.loss_scale = float(loss_scale)
        dev = store.grad.device
        self.ctrl = torch.zeros(3, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(4, dtype=torch.float32, device=dev)

    def compute(self, grad_scale):
        """grad_scale: the optimizer's own multiplier (1/world, 1/accum_steps); the norm is that of the gradient the update sees."""
        lib = _lib.load()
        _lib.check(lib.passl_b200_grad_norm_finite(self.store.grad.data_ptr(), self.store.numel, float(grad_scale) / self.loss_scale,
                                                   self.clip_norm, self.clip_norm_max, int(self.always_clip), self.ctrl.data_ptr(),
                                                   self._scratch.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "grad_norm_finite")
        return self.ctrl

    @property
    def global_norm(self):
        return self.ctrl[2]

    @property
    def found_inf(self):
        return self.ctrl[1]


class _FlatOptimizer:
    def __init__(self, store, lr, grad_clip=None):
        self.store = store
        self.lr = float(lr)
        self.grad_scale = 1.0 / get_world_size()     # the mean of sync_utils.py:41
        self._step = 0
        self.grad_control = None
        if grad_clip:
            self.set_grad_clip(grad_clip)

    def set_grad_clip(self, cfg):
        """cfg: dict(name='ClipGradByGlobalNorm', clip_norm=..., clip_norm_max=..., always_clip=...) as in the reference YAMLs
        (Optimizer.grad_clip), a GradControl, or None."""
        if cfg is None or isinstance(cfg, GradControl):
            self.grad_control = cfg
            return
        cfg = dict(cfg)
        name = cfg.pop("name", "ClipGradByGlobalNorm")
        if name != "ClipGradByGlobalNorm":
            raise NotImplementedError("grad_clip %r is not built (ClipGradByGlobalNorm is)" % name)
        cfg.pop("no_clip_list", None)
        self.grad_control = GradControl(self.store, **cfg)

    def _ctrl_ptr(self):
        """runs the one-pass norm / finite check when a GradControl is attached; returns the device control word (or 0)"""
        if self.grad_control is None:
            return 0
        # the norm is that of the gradient the update sees (mean over ranks / accumulation steps, unscaled): the control
        # multiplier carries grad_scale itself, so the kernels are then launched with a unit host-side scale (_gscale)
        self.grad_control.compute(self.grad_scale)
        return self.grad_control.ctrl.data_ptr()

    def _gscale(self, ctrl):
        return 1.0 if ctrl else self.grad_scale

    def clear_grad(self):
        self.store.zero_grad()

    zero_grad = clear_grad

    def set_lr(self, lr):
        self.lr = float(lr)

    def set_state_dict(self, state):
        """Inverse of state_dict() (paddle optimizer API name): restores the flat moment buffers, step counter and lr."""
        for k, v in state.items():
            cur = getattr(self, "_step" if k == "step" else k, None)
            if torch.is_tensor(cur) and torch.is_tensor(v):
                cur.copy_(v.to(cur.device))
            elif k == "step":
                self._step = int(v)
            elif k == "lr":
                self.lr = float(v)

    load_state_dict = set_state_dict

    def get_lr(self):
        return self.lr


class Momentum(_FlatOptimizer):
    """passl/optimizer/momentum.py:60-158 (L2 decay folded into the gradient, no nesterov)."""

    def __init__(self, store, lr=0.01, momentum=0.9, weight_decay=0.0):
        super().__init__(store, lr)
        self.momentum, self.weight_decay = momentum, weight_decay
        self.velocity = torch.zeros_like(store.master)
        self._ranges = trainable_ranges(store)       # the scalar-decay kernel has no per-tensor table: skip frozen runs on the host

    def step(self):
        lib, s = _lib.load(), self.store
        st = torch.cuda.current_stream().cuda_stream
        ctrl = self._ctrl_ptr()
        for off, n in self._ranges:                  # one launch when nothing is frozen (the usual case)
            _lib.check(lib.passl_b200_sgd_momentum(s.master.data_ptr() + 4 * off, s.grad.data_ptr() + 4 * off,
                                                   self.velocity.data_ptr() + 4 * off, s.bf16.data_ptr() + 2 * off, self.lr,
                                                   self.momentum, self.weight_decay, self._gscale(ctrl), ctrl, n, st), "sgd_momentum")
        self._step += 1

    def state_dict(self):
        return dict(velocity=self.velocity, step=self._step, lr=self.lr)


class LarsMomentumOptimizer(_FlatOptimizer):
    """paddle.fluid LarsMomentumOptimizer as configured by passl_v110/solver/optimizer.py:25 and the SimCLR YAML
    (configs/simclr/simclr_r50_IM.yaml:116-120); update rule of passl/optimizer/momentum_lars.py:56-114.  Tensors whose decay is
    zero take the plain momentum step (no trust ratio), like the lars_momentum op."""

    def __init__(self, store, lr=0.1, momentum=0.9, lars_weight_decay=1e-4, lars_coeff=0.001, epsilon=0.0,
                 exclude_from_weight_decay=None, exclude_structured=None):
        super().__init__(store, lr)
        self.momentum, self.coeff, self.eps = momentum, lars_coeff, epsilon
        self.velocity = torch.zeros_like(store.master)
        self.seg_wd = _wd_table(store, lars_weight_decay, exclude_from_weight_decay, exclude_structured)
        self.norms = torch.zeros(2 * len(store.params), dtype=torch.float32, device=store.master.device)

    def step(self):
        lib, s = _lib.load(), self.store
        ctrl = self._ctrl_ptr()
        _lib.check(lib.passl_b200_lars_momentum(s.master.data_ptr(), s.grad.data_ptr(), self.velocity.data_ptr(),
                                                s.bf16.data_ptr(), s.block_seg.data_ptr(), self.seg_wd.data_ptr(),
                                                self.norms.data_ptr(), len(s.params), self.lr, self.momentum, self.coeff,
                                                self.eps, self._gscale(ctrl), ctrl, s.numel,
                                                torch.cuda.current_stream().cuda_stream), "lars_momentum")
        self._step += 1

    def state_dict(self):
        return dict(velocity=self.velocity, step=self._step, lr=self.lr)


class AdamW(_FlatOptimizer):
    """Decoupled-decay Adam with bias correction (the `adamw` op behind passl/optimizer/adamw.py:52-138 and paddle.optimizer.AdamW).

    Which tensors decay follows the reference's three AdamW call paths:
      * v2.5 `passl.optimizer.AdamW` (MoCo v3 YAML) and v110 `paddle.optimizer.AdamW` (CLIP YAML): EVERY trainable tensor decays,
        1-d ones included, unless its state_dict name contains one of the substrings in `no_weight_decay_name`
        (passl/optimizer/utils/group_params.py:175) / `exclude_from_weight_decay` (passl_v110/solver/builder.py:204-214);
      * MAE pre-training builds its groups with add_weight_decay: 1-d tensors and `*.bias` undecayed
        (tasks/ssl/mae/util/optim_factory.py:21-38) -> `one_dim_no_decay=True`;
      * frozen tensors are never in a group.
    `no_decay` takes regexes on the state_dict name for callers who prefer them; `betas` / `eps` / `epsilon` cover both YAML
    spellings; `use_master_param` / `exp_avg_force_fp32` are accepted and moot (master weights and moments are always fp32 here)."""

    def __init__(self, store, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01, no_decay=None, lr_ratio=None,
                 epsilon=None, betas=None, one_dim_no_decay=False, exclude_from_weight_decay=None, no_weight_decay_name=None,
                 use_master_param=True, exp_avg_force_fp32=True):
        super().__init__(store, lr)
        if betas is not None:
            if isinstance(betas, str):                   # the YAML spells it "(0.9, 0.999)"
                import ast
                betas = ast.literal_eval(betas)
            beta1, beta2 = (float(b) for b in betas)
        self.beta1, self.beta2, self.eps = beta1, beta2, float(eps if epsilon is None else epsilon)
        self.m = torch.zeros_like(store.master)
        self.v = torch.zeros_like(store.master)
        pats = [re.compile(p) for p in (no_decay or [])]
        subs = list(exclude_from_weight_decay or []) + list(no_weight_decay_name or [])

        def decay(n, p):
            if not p.requires_grad:
                return 0.0
            if one_dim_no_decay and (p.dim() <= 1 or n.endswith(".bias")):
                return 0.0
            if any(s_ in n for s_ in subs) or any(r.search(n) for r in pats):
                return 0.0
            return weight_decay
        self.seg_wd = store.segment_values(decay)
        self.seg_lr = store.segment_values(lr_ratio) if lr_ratio is not None else None

    def step(self):
        lib, s = _lib.load(), self.store
        self._step += 1
        ctrl = self._ctrl_ptr()
        _lib.check(lib.passl_b200_adamw(s.master.data_ptr(), s.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                        s.bf16.data_ptr(), s.block_seg.data_ptr(), self.seg_wd.data_ptr(),
                                        self.seg_lr.data_ptr() if self.seg_lr is not None else 0, self.lr, self.beta1,
                                        self.beta2, self.eps, self._step, self._gscale(ctrl), ctrl, s.numel,
                                        torch.cuda.current_stream().cuda_stream), "adamw")

    def state_dict(self):
        return dict(m=self.m, v=self.v, step=self._step, lr=self.lr)


def build_optimizer(cfg, store):
    cfg = dict(cfg)
    name = cfg.pop("name")
    cls = {"Momentum": Momentum, "LarsMomentumOptimizer": LarsMomentumOptimizer, "AdamW": AdamW}[name]
    grad_clip = cfg.pop("grad_clip", None)          # Optimizer.grad_clip of the v2.5 YAMLs (passl/optimizer/__init__.py:64-70)
    opt = cls(store, **cfg)
    if grad_clip:
        opt.set_grad_clip(grad_clip)
    return opt
