"""passl_b200/optimizer/lr.py against curves produced by the reference's own scheduler classes and builders
(tests/golden/make_golden_lr.py -> reference_lr.npz): every value, driven with the protocol of the respective reference trainer."""
import os

import numpy as np
import pytest

from passl_b200.optimizer import lr as L

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_lr.npz"))
TOL = dict(rtol=1e-12, atol=1e-15)


def _curve(s, n):
    out = []
    for _ in range(n):
        out.append(s())
        s.step()
    return np.array(out)


@pytest.mark.parametrize("tag,scaling", [("simclr_sqrt", "sqrt"), ("simclr_linear", "linear")])
def test_simclr_recipe(tag, scaling):
    total_images, per_gpu, epochs, wu, end_lr = G[tag + "_args"]
    total_images, per_gpu, epochs, wu = int(total_images), int(per_gpu), int(epochs), int(wu)
    cfg = dict(name="simclrCosineWarmup", learning_rate_scaling=scaling, total_images=total_images, warmup_epochs=wu, start_lr=0,
               end_lr=float(end_lr), T_max=200)
    s = L.build_lr_scheduler_simclr(cfg, total_images // (per_gpu * 8), per_gpu * 8, epochs, 0)
    want = G[tag]
    np.testing.assert_allclose(_curve(s, len(want)), want, **TOL)
    peak = end_lr * (np.sqrt(per_gpu * 8) if scaling == "sqrt" else per_gpu * 8 / 256.0)
    assert want[0] == 0.0 and abs(want.max() - peak) < 1e-12 and want[-1] < 1e-3 * peak          # warm-up from 0, cosine to ~0


def test_moco_clip_multistep_vit():
    np.testing.assert_allclose(_curve(L.build_lr_scheduler(dict(name="CosineAnnealingDecay", learning_rate=0.03, T_max=5), 13), 66),
                               G["moco_cosine"], rtol=1e-9, atol=1e-15)           # the reference steps paddle's recurrence: rounding only
    s = L.build_lr_scheduler(dict(name="LinearWarmup", learning_rate=dict(name="CosineAnnealingDecay", learning_rate=1e-4, T_max=10, eta_min=1e-6),
                                  warmup_steps=5, start_lr=0, end_lr=1e-4), 7)
    np.testing.assert_allclose(_curve(s, 106), G["clip_warmup_cosine"], **TOL)
    np.testing.assert_allclose(_curve(L.build_lr_scheduler(dict(name="MultiStepDecay", learning_rate=0.1, milestones=[2, 4], gamma=0.1), 5), 30),
                               G["multistep"], **TOL)
    np.testing.assert_allclose(_curve(L.ViTLRScheduler(3e-3, 60, decay_type="cosine", warmup_steps=9), 70), G["vit_cosine"], **TOL)
    np.testing.assert_allclose(_curve(L.ViTLRScheduler(3e-3, 60, decay_type="linear", warmup_steps=0), 70), G["vit_linear"], **TOL)
    with pytest.raises(NotImplementedError):
        L.build_lr_scheduler(dict(name="NoSuchSchedule"), 1)


@pytest.mark.parametrize("tag,kw", [("timm_step_prefix", dict(decay_unit="step", warmup_epoch=2, warmup_prefix=True, eta_min=0.0)),
                                    ("timm_step", dict(decay_unit="step", warmup_epoch=1, warmup_prefix=False, eta_min=1e-5))])
def test_timm_cosine_v2_protocol(tag, kw):
    """v2.5: the optimizer reads get_lr() and the loop then calls step(global_step) (optimizer.py:117-123,216-222)."""
    s = L.build_lr_scheduler_v2(dict(name="TimmCosine", learning_rate=0.0024, warmup_start_lr=0.0, **kw), epochs=6, step_each_epoch=11)
    vals = []
    for global_step in range(1, 67):
        vals.append(s.get_lr())
        s.step(global_step)
    np.testing.assert_allclose(vals, G[tag], **TOL)
    np.testing.assert_allclose([s.lr_at(k - 1) for k in range(1, 67)], G[tag], **TOL)            # step k runs at lr_at(k - 1)


def test_mae_half_cycle():
    s = L.MAEHalfCycleCosine(2.4e-3, 1e-6, warmup_epochs=2, epochs=8, step_each_epoch=9)
    np.testing.assert_allclose(_curve(s, 72), G["mae_half_cycle"], **TOL)


def test_state_roundtrip_and_explicit_index():
    a = L.build_lr_scheduler(dict(name="CosineAnnealingDecay", learning_rate=0.03, T_max=5), 13)
    for _ in range(17):
        a.step()
    b = L.build_lr_scheduler(dict(name="CosineAnnealingDecay", learning_rate=0.03, T_max=5), 13)
    b.set_state_dict(a.state_dict())
    assert b() == a() and b.last_epoch == 17 and b.step() == a.step()
    assert a.step(3) == a.lr_at(3)
