"""Golden learning-rate curves produced by the reference's OWN scheduler classes and builders
(passl_v110/solver/lr_scheduler.py, passl_v110/solver/builder.py, passl/scheduler/lr_scheduler.py, tasks/ssl/mae/util/lr_sched.py),
executed over a stand-in for `paddle.optimizer.lr` — TEST INFRASTRUCTURE, run in the build container only:

    cd tests/golden && python make_golden_lr.py        ->  reference_lr.npz

The stand-in restates the documented stepping protocol of paddle 2.4's LRScheduler / LinearWarmup / CosineAnnealingDecay /
MultiStepDecay (constructor performs one step(); step(epoch) jumps and prefers _get_closed_form_lr); everything specific to the
reference (Cosinesimclr, simclrCosineWarmup, ViTLRScheduler, TimmCosine, the builders' epoch->iteration conversions and the SimCLR
batch-size scaling, MAE's adjust_learning_rate) is the reference's code.  Curves are sampled the way the reference trainers drive
the objects: v110 `scheduler.step()` after every iteration (hooks/lr_scheduler_hook.py:28), v2.5 `step(global_step)` after every
optimizer step (optimizer.py:216-222), MAE a call per iteration with the fractional epoch."""
import importlib
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class LRScheduler:
    def __init__(self, learning_rate=0.1, last_epoch=-1, verbose=False):
        self.base_lr, self.last_lr, self.last_epoch, self.verbose = float(learning_rate), float(learning_rate), last_epoch, verbose
        self.step()

    def __call__(self):
        return self.last_lr

    def step(self, epoch=None):
        if epoch is None:
            self.last_epoch += 1
            self.last_lr = self.get_lr()
        else:
            self.last_epoch = epoch
            self.last_lr = self._get_closed_form_lr() if hasattr(self, "_get_closed_form_lr") else self.get_lr()


class CosineAnnealingDecay(LRScheduler):
    def __init__(self, learning_rate, T_max, eta_min=0, last_epoch=-1, verbose=False):
        self.T_max, self.eta_min = T_max, float(eta_min)
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        if self.last_epoch == 0:
            return self.base_lr
        if (self.last_epoch - 1 - self.T_max) % (2 * self.T_max) == 0:
            return self.last_lr + (self.base_lr - self.eta_min) * (1 - math.cos(math.pi / self.T_max)) / 2
        return (1 + math.cos(math.pi * self.last_epoch / self.T_max)) / (1 + math.cos(math.pi * (self.last_epoch - 1) / self.T_max)) * (
            self.last_lr - self.eta_min) + self.eta_min

    def _get_closed_form_lr(self):
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * self.last_epoch / self.T_max)) / 2


class MultiStepDecay(LRScheduler):
    def __init__(self, learning_rate, milestones, gamma=0.1, last_epoch=-1, verbose=False):
        self.milestones, self.gamma = milestones, gamma
        super().__init__(learning_rate, last_epoch, verbose)

    def get_lr(self):
        for i in range(len(self.milestones)):
            if self.last_epoch < self.milestones[i]:
                return self.base_lr * (self.gamma ** i)
        return self.base_lr * (self.gamma ** len(self.milestones))


class LinearWarmup(LRScheduler):
    def __init__(self, learning_rate, warmup_steps, start_lr, end_lr, last_epoch=-1, verbose=False):
        self.learning_rate, self.warmup_steps, self.start_lr, self.end_lr = learning_rate, warmup_steps, start_lr, end_lr
        assert end_lr > start_lr
        super().__init__(start_lr, last_epoch, verbose)

    def get_lr(self):
        if self.last_epoch < self.warmup_steps:
            return (self.end_lr - self.start_lr) * float(self.last_epoch) / float(self.warmup_steps) + self.start_lr
        if isinstance(self.learning_rate, LRScheduler):
            self.learning_rate.step(self.last_epoch - self.warmup_steps)
            return self.learning_rate()
        return self.learning_rate


def install():
    sys.path.insert(0, HERE)
    import make_golden
    import paddle_shim
    make_golden.setup()                      # paddle stand-in + the reference packages exposed without running their __init__
    paddle = sys.modules["paddle"]
    lr = types.ModuleType("paddle.optimizer.lr")
    lr.LRScheduler, lr.CosineAnnealingDecay, lr.MultiStepDecay, lr.LinearWarmup = LRScheduler, CosineAnnealingDecay, MultiStepDecay, LinearWarmup
    opt = paddle_shim._Permissive("paddle.optimizer")
    opt.lr = lr
    paddle.optimizer = opt
    sys.modules["paddle.optimizer"], sys.modules["paddle.optimizer.lr"] = opt, lr
    clip = paddle_shim._Permissive("paddle.nn.clip")          # builder.py imports the gradient-clip classes at module scope
    sys.modules["paddle.nn.clip"] = clip
    sys.modules["paddle.nn"].clip = clip
    paddle_shim.fake_package("passl_v110.solver", os.path.join(REF, "passl_v110", "solver"))
    paddle_shim.fake_package("passl.scheduler", os.path.join(REF, "passl", "scheduler"))
    utils = paddle_shim.fake_package("passl.utils", os.path.join(REF, "passl", "utils"))
    utils.logger = types.SimpleNamespace(warning=lambda *a, **k: None, debug=lambda *a, **k: None, info=lambda *a, **k: None)
    return paddle


class Cfg(dict):
    """attribute-style dict like the reference's config nodes"""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def curve_v110(s, n):
    out = []
    for _ in range(n):
        out.append(s())
        s.step()
    return np.array(out, dtype=np.float64)


def main():
    install()
    # the registry / builder modules import the whole solver package; load the two files the path needs
    builder = importlib.import_module("passl_v110.solver.builder")
    sched = importlib.import_module("passl_v110.solver.lr_scheduler")
    out = {}
    # SimCLR headline recipe (configs/simclr/simclr_r50_IM.yaml:106-113) at a reduced image count so the curve is short
    for tag, (scaling, total_images, per_gpu, epochs, wu) in {"simclr_sqrt": ("sqrt", 4096, 32, 20, 2), "simclr_linear": ("linear", 5000, 16, 7, 1)}.items():
        cfg = Cfg(name="simclrCosineWarmup", learning_rate_scaling=scaling, total_images=total_images, warmup_epochs=wu, start_lr=0,
                  end_lr=1.0 if scaling == "sqrt" else 0.3, T_max=200)
        s = builder.build_lr_scheduler_simclr(cfg, total_images // (per_gpu * 8), per_gpu * 8, epochs, 0)
        n = total_images * epochs // (per_gpu * 8) + 1
        out[tag] = curve_v110(s, n)
        out[tag + "_args"] = np.array([total_images, per_gpu, epochs, wu, cfg["end_lr"]], dtype=np.float64)
    # MoCo v2 (configs/moco/moco_v2_r50.yaml:84-87): CosineAnnealingDecay, T_max in epochs
    s = builder.build_lr_scheduler(Cfg(name="CosineAnnealingDecay", learning_rate=0.03, T_max=5), 13)
    out["moco_cosine"] = curve_v110(s, 5 * 13 + 1)
    # CLIP (configs/clip/vit-b-32.yaml:51-60): LinearWarmup around CosineAnnealingDecay(eta_min)
    s = builder.build_lr_scheduler(Cfg(name="LinearWarmup", learning_rate=Cfg(name="CosineAnnealingDecay", learning_rate=1e-4, T_max=10, eta_min=1e-6),
                                       warmup_steps=5, start_lr=0, end_lr=1e-4), 7)
    out["clip_warmup_cosine"] = curve_v110(s, 15 * 7 + 1)
    s = builder.build_lr_scheduler(Cfg(name="MultiStepDecay", learning_rate=0.1, milestones=[2, 4], gamma=0.1), 5)
    out["multistep"] = curve_v110(s, 30)
    for tag, kw in {"vit_cosine": dict(decay_type="cosine", warmup_steps=9), "vit_linear": dict(decay_type="linear", warmup_steps=0)}.items():
        s = sched.ViTLRScheduler(learning_rate=3e-3, T_max=60, **kw)
        out[tag] = curve_v110(s, 70)
    # v2.5 TimmCosine as the MoCo v3 pre-training YAML configures it; driven by step(global_step) after each optimizer step
    v2 = importlib.import_module("passl.scheduler.lr_scheduler")
    for tag, kw in {"timm_step_prefix": dict(decay_unit="step", warmup_epoch=2, warmup_prefix=True, eta_min=0.0),
                    "timm_step": dict(decay_unit="step", warmup_epoch=1, warmup_prefix=False, eta_min=1e-5)}.items():
        s = v2.TimmCosine(learning_rate=0.0024, step_each_epoch=11, epochs=6, warmup_start_lr=0.0, **kw)
        vals = []
        for global_step in range(1, 6 * 11 + 1):
            vals.append(s.get_lr())                 # what optimizer step `global_step` reads (optimizer.py:117-123)
            s.step(global_step)
        out[tag] = np.array(vals, dtype=np.float64)
    # MAE (tasks/ssl/mae/util/lr_sched.py): per-iteration call with the fractional epoch
    mae = importlib.import_module("tasks.ssl.mae.util.lr_sched") if os.path.isfile(os.path.join(REF, "tasks/__init__.py")) else None
    if mae is None:
        spec = importlib.util.spec_from_file_location("mae_lr_sched", os.path.join(REF, "tasks/ssl/mae/util/lr_sched.py"))
        mae = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mae)
    args = types.SimpleNamespace(lr=2.4e-3, min_lr=1e-6, warmup_epochs=2, epochs=8)
    opt = types.SimpleNamespace(param_groups=[{}])
    steps = 9
    out["mae_half_cycle"] = np.array([mae.adjust_learning_rate(opt, i / steps + e, args) for e in range(8) for i in range(steps)])
    np.savez_compressed(os.path.join(HERE, "reference_lr.npz"), **out)
    print("wrote reference_lr.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    import importlib.util
    main()
