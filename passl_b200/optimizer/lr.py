"""Learning-rate schedules of the hot-path configs, as pure functions of the step index.

The reference drives `paddle.optimizer.lr.LRScheduler` objects: the constructor performs one `step()` (last_epoch -1 -> 0),
`step()` advances by one or jumps to an explicit index, `scheduler()` returns the current value, and the optimizer reads it when it
steps.  Here every schedule implements `lr_at(t)` — the rate of iteration t, t = 0 for the first optimizer step — and the base class
supplies that stepping protocol on top, so nested schedules (warm-up around a decay) need no hidden state.

    config                         schedule                                                     reference
    moco_v2_r50.yaml               CosineAnnealingDecay(lr, T_max epochs * iters_per_epoch)      paddle.optimizer.lr; solver/builder.py:28-30
    simclr_r50_IM.yaml             simclrCosineWarmup: linear 0 -> lr over the warm-up steps,    passl_v110/solver/lr_scheduler.py:106-139,
                                   then lr (1 + cos(pi t' / T_max)) / 2                          solver/builder.py:46-66, engine/trainer.py:157-163
    clip/vit-b-32.yaml             LinearWarmup(CosineAnnealingDecay(eta_min))                   solver/builder.py:34-38
    mocov3 ..pt_in1k.. yaml        TimmCosine(decay_unit=step, warmup_prefix)                    passl/scheduler/lr_scheduler.py:22-77
    mae pretrain                   half-cycle cosine after linear warm-up, per iteration         tasks/ssl/mae/util/lr_sched.py:21-30
    ViT fine-tune recipes          ViTLRScheduler (cosine / linear with multiplicative warm-up)  passl_v110/solver/lr_scheduler.py:142-180

paddle's CosineAnnealingDecay advances by a recurrence on the previous value when stepped one by one and by the closed form when
given an index; the two agree to rounding, and the closed form is what is implemented here.
"""
import bisect
import math


class LRScheduler:
    """Stepping protocol of paddle.optimizer.lr.LRScheduler over `lr_at`."""

    def __init__(self, learning_rate, last_epoch=-1):
        self.base_lr = float(learning_rate)
        self.last_epoch = last_epoch
        self.last_lr = self.base_lr
        self.step()

    def lr_at(self, t):
        raise NotImplementedError

    def get_lr(self):
        return self.lr_at(self.last_epoch)

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        self.last_lr = self.get_lr()
        return self.last_lr

    def __call__(self):
        return self.last_lr

    def state_dict(self):
        return {"last_epoch": self.last_epoch, "last_lr": self.last_lr}

    def set_state_dict(self, state):
        self.last_epoch = int(state["last_epoch"])
        self.last_lr = self.get_lr()


def _rate(lr, t):
    return lr.lr_at(t) if isinstance(lr, LRScheduler) else float(lr)


class CosineAnnealingDecay(LRScheduler):
    """eta_min + (lr - eta_min) (1 + cos(pi t / T_max)) / 2  (configs/moco/moco_v2_r50.yaml:84-87)."""

    def __init__(self, learning_rate, T_max, eta_min=0.0, last_epoch=-1, **kwargs):
        self.T_max, self.eta_min = T_max, float(eta_min)
        super().__init__(learning_rate, last_epoch)

    def lr_at(self, t):
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t / self.T_max)) / 2


class MultiStepDecay(LRScheduler):
    """lr * gamma^(number of milestones <= t)."""

    def __init__(self, learning_rate, milestones, gamma=0.1, last_epoch=-1, **kwargs):
        self.milestones, self.gamma = sorted(milestones), gamma
        super().__init__(learning_rate, last_epoch)

    def lr_at(self, t):
        return self.base_lr * self.gamma ** bisect.bisect_right(self.milestones, t)


class LinearWarmup(LRScheduler):
    """start_lr -> end_lr linearly over `warmup_steps`, then the wrapped rate / schedule re-indexed from 0."""

    def __init__(self, learning_rate, warmup_steps, start_lr, end_lr, last_epoch=-1, **kwargs):
        assert end_lr > start_lr, "end_lr {} must be greater than start_lr {}".format(end_lr, start_lr)
        self.learning_rate, self.warmup_steps, self.start_lr, self.end_lr = learning_rate, warmup_steps, float(start_lr), float(end_lr)
        super().__init__(start_lr, last_epoch)

    def lr_at(self, t):
        if t < self.warmup_steps:
            return (self.end_lr - self.start_lr) * float(t) / float(self.warmup_steps) + self.start_lr
        return _rate(self.learning_rate, t - self.warmup_steps)


class Cosinesimclr(LRScheduler):
    """lr (1 + cos(pi t / T_max)) / 2  (solver/lr_scheduler.py:106-114)."""

    def __init__(self, learning_rate, T_max, last_epoch=-1, **kwargs):
        self.T_max = T_max
        super().__init__(learning_rate, last_epoch)

    def lr_at(self, t):
        return self.base_lr * (1 + math.cos(math.pi * t / self.T_max)) / 2


class simclrCosineWarmup(LinearWarmup):
    """SimCLR recipe: warm up from 0 to `lr`, then Cosinesimclr over T_max steps (solver/lr_scheduler.py:117-139).  Built by
    build_lr_scheduler_simclr, which derives lr, warmup_steps and T_max from the batch size and the image count."""

    def __init__(self, lr, warmup_steps, T_max, current_iter=0, last_epoch=-1, **kwargs):
        super().__init__(Cosinesimclr(lr, T_max), warmup_steps, 0.0, lr, last_epoch)


class ViTLRScheduler(LRScheduler):
    """cosine / linear decay over (t - warmup) / (T_max - warmup) times min(1, t / warmup) (solver/lr_scheduler.py:142-180)."""

    def __init__(self, learning_rate, T_max, decay_type="cosine", linear_end=1e-5, warmup_steps=0, last_epoch=-1, **kwargs):
        self.T_max, self.decay_type, self.linear_end = T_max, decay_type, linear_end
        self.warmup_steps = min(warmup_steps, T_max)
        super().__init__(learning_rate, last_epoch)

    def lr_at(self, t):
        progress = min(1.0, max(0.0, (t - self.warmup_steps) / float(self.T_max - self.warmup_steps)))
        if self.decay_type == "linear":
            lr = self.linear_end + (self.base_lr - self.linear_end) * (1.0 - progress)
        else:
            lr = 0.5 * self.base_lr * (1.0 + math.cos(math.pi * progress))
        if self.warmup_steps:
            lr = lr * min(1.0, t / self.warmup_steps)
        return lr


class TimmCosine(LRScheduler):
    """passl/scheduler/lr_scheduler.py:22-77 (MoCo v3 pre-training: decay_unit=step, warmup_epoch=40, warmup_prefix).  The
    reference object starts at last_epoch = -1 and is moved with step(global_step) after every optimizer step, so optimizer step
    k (1-based) runs at lr_at(k - 1); negative indices clamp to the warm-up start."""

    def __init__(self, learning_rate, step_each_epoch, epochs, decay_unit="epoch", eta_min=0.0, warmup_epoch=0, warmup_start_lr=0.0,
                 warmup_prefix=False, last_epoch=-1, **kwargs):
        assert decay_unit in ("step", "epoch")
        warmup_epoch = min(warmup_epoch, epochs)
        if decay_unit == "step":
            self.T_max, self.warmup_steps = epochs * step_each_epoch, int(round(warmup_epoch * step_each_epoch))
        else:
            self.T_max, self.warmup_steps = epochs, warmup_epoch
        self.decay_unit, self.eta_min, self.warmup_start_lr, self.warmup_prefix = decay_unit, eta_min, warmup_start_lr, warmup_prefix
        self.base_lr = self.last_lr = float(learning_rate)          # no constructor step() in the reference class
        self.last_epoch = last_epoch

    def lr_at(self, t):
        if t < self.warmup_steps:
            return float(max(0, t)) * (self.base_lr - self.warmup_start_lr) / float(self.warmup_steps) + self.warmup_start_lr
        T = self.T_max
        if self.warmup_prefix:
            t, T = t - self.warmup_steps, self.T_max - self.warmup_steps
        cur = t - self.T_max * (t // self.T_max)
        return self.eta_min + 0.5 * (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * cur / T))


class MAEHalfCycleCosine(LRScheduler):
    """tasks/ssl/mae/util/lr_sched.py:21-30, called once per iteration with the fractional epoch
    `data_iter_step / len(data_loader) + epoch` (engine_pretrain.py:64-67): t counts iterations here."""

    def __init__(self, lr, min_lr, warmup_epochs, epochs, step_each_epoch, last_epoch=-1, **kwargs):
        self.min_lr, self.warmup_epochs, self.epochs, self.step_each_epoch = min_lr, warmup_epochs, epochs, step_each_epoch
        super().__init__(lr, last_epoch)

    def lr_at(self, t):
        epoch = t / self.step_each_epoch
        if epoch < self.warmup_epochs:
            return self.base_lr * epoch / self.warmup_epochs
        return self.min_lr + (self.base_lr - self.min_lr) * 0.5 * (
            1.0 + math.cos(math.pi * (epoch - self.warmup_epochs) / (self.epochs - self.warmup_epochs)))


_REGISTRY = {c.__name__: c for c in (CosineAnnealingDecay, MultiStepDecay, LinearWarmup, Cosinesimclr, simclrCosineWarmup,
                                     ViTLRScheduler, TimmCosine, MAEHalfCycleCosine)}


def build_lr_scheduler(cfg, iters_per_epoch):
    """passl_v110/solver/builder.py:26-42: epoch-denominated keys of the YAML are converted to iterations here."""
    cfg = dict(cfg)
    name = cfg.pop("name")
    if name in ("CosineAnnealingDecay", "ViTLRScheduler"):
        cfg["T_max"] = cfg["T_max"] * iters_per_epoch
    elif name == "MultiStepDecay":
        cfg["milestones"] = [x * iters_per_epoch for x in cfg["milestones"]]
    elif name == "LinearWarmup":
        if isinstance(cfg["learning_rate"], dict) or hasattr(cfg["learning_rate"], "keys"):
            cfg["learning_rate"] = build_lr_scheduler(cfg["learning_rate"], iters_per_epoch)
        cfg["warmup_steps"] = cfg["warmup_steps"] * iters_per_epoch
    elif name not in _REGISTRY:
        raise NotImplementedError(name)
    return _REGISTRY[name](**cfg)


def build_lr_scheduler_simclr(cfg, iters_per_epoch, batch_size, epochs, current_iter=0):
    """passl_v110/solver/builder.py:46-66.  `batch_size` is the global batch the recipe assumes (the reference trainer passes
    per-GPU batch * 8, engine/trainer.py:161-163)."""
    cfg = dict(cfg)
    name = cfg["name"]
    if name != "simclrCosineWarmup":
        return build_lr_scheduler(cfg, iters_per_epoch)
    warmup_steps = int(round(cfg["warmup_epochs"] * cfg["total_images"] // batch_size))
    total_steps = cfg["total_images"] * epochs // batch_size + 1
    scaling = cfg.get("learning_rate_scaling", "linear")
    if scaling == "linear":
        lr = cfg["end_lr"] * batch_size / 256.0
    elif scaling == "sqrt":
        lr = cfg["end_lr"] * math.sqrt(batch_size)
    else:
        raise ValueError("learning_rate_scaling must be linear or sqrt, got %r" % (scaling,))
    return simclrCosineWarmup(lr, warmup_steps, total_steps - warmup_steps, current_iter)


def build_lr_scheduler_v2(lr_config, epochs, step_each_epoch):
    """passl/scheduler/__init__.py:22-36 (the `LRScheduler:` section of the v2.5 YAMLs)."""
    cfg = dict(lr_config)
    if "name" not in cfg:
        return cfg["learning_rate"]
    name = cfg.pop("name")
    if name not in _REGISTRY:
        raise NotImplementedError(name)
    return _REGISTRY[name](epochs=epochs, step_each_epoch=step_each_epoch, **cfg)
