"""The v2.5 `Engine` surface end to end on the GPU (engine/engine.py): MoCo v3 YAML -> model, fused AdamW, TimmCosine schedule, epoch
loop, checkpoints, resume.  Same small ViT dimensions as tests/test_models_gpu.py::test_mocov3_small_step."""
import functools
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CFG = os.path.join(os.path.dirname(HERE), "configs/mocov3/mocov3_vit_base_patch16_224_pt.yaml")


def _small_factory(**kw):
    from passl_b200.models.mocov3 import MoCoV3Pretrain, MoCoV3ViT
    enc = functools.partial(MoCoV3ViT, img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, qkv_bias=True, epsilon=1e-6,
                            stop_grad_conv1=True)
    return MoCoV3Pretrain(enc, dim=128, mlp_dim=256, T=0.2, **kw)


def test_engine_trains_saves_and_resumes(monkeypatch, tmp_path, capsys):
    import passl_b200.models as M
    from passl_b200.engine.engine import Engine, SyntheticTwoViewLists
    from passl_b200.utils.config import get_config
    monkeypatch.setattr(M, "small_mocov3_pretrain", _small_factory, raising=False)
    over = ["Model.name=small_mocov3_pretrain", "Global.epochs=2", "Global.max_train_step=null", "Global.output_dir=%s" % tmp_path,
            "Global.print_batch_step=2", "LRScheduler.warmup_epoch=1", "LRScheduler.learning_rate=0.001",
            "DataLoader.Train.sampler.batch_size=16"]
    dev = torch.device("cuda", 0)

    def make():
        cfg = get_config(CFG, over)
        cfg["Global"]["max_train_step"] = None
        return Engine(cfg, dataloader=SyntheticTwoViewLists(16, 3, dev, size=64))
    e = make()
    frozen = e.model.base_encoder.vit.pos_embed.detach().clone()
    w0 = e.store.master.clone()
    assert e.train() == 6
    log = capsys.readouterr().out
    assert "[Train][Epoch 1/2][Iter: 2/3]" in log and "[Train][Epoch 2/2][Iter: 3/3]" in log
    assert torch.isfinite(e.store.master).all() and not torch.equal(w0, e.store.master)
    assert torch.equal(frozen, e.model.base_encoder.vit.pos_embed.detach())
    assert e.lr_scheduler.last_epoch == 6 and abs(e.optimizer.lr - e.lr_scheduler.lr_at(5)) < 1e-12      # last step ran at lr_at(k - 1)
    base = os.path.join(str(tmp_path), "small_mocov3_pretrain")
    assert {"epoch_1.pdparams", "epoch_2.pdparams", "epoch_2_base_encoder.pdparams", "epoch_2.pdstates", "epoch_2.opt.pt"} <= set(os.listdir(base))
    # resume epoch 1 in a fresh engine and run epoch 2: same data, same schedule -> weights close to the uninterrupted run
    e2 = make()
    e2.resume(os.path.join(base, "epoch_1"))
    assert (e2.cur_epoch_id, e2.global_step) == (1, 3) and e2.model.steps == 3
    assert e2.train() == 6
    a, b = e.store.master, e2.store.master                         # split-K atomics make single elements wander by ~lr: compare in norm
    assert ((a - b).norm() / a.norm()).item() < 1e-2
    assert np.isfinite(float(e2.store.master.sum().item()))


def test_engine_gradient_accumulation_and_grad_clip(monkeypatch, tmp_path):
    """Global.accum_steps (contrastive_learning_loop.py:31-63) and Optimizer.grad_clip (ClipGradByGlobalNorm, grad_clip.py:30-84):
    a step over a 32-sample batch with accum_steps=2 equals two forward/backward passes over its halves, gradients summed, one
    update with the mean gradient clipped to the global norm."""
    import passl_b200.models as M
    from passl_b200.engine.engine import Engine, SyntheticTwoViewLists
    from passl_b200.utils.config import get_config
    monkeypatch.setattr(M, "small_mocov3_pretrain", _small_factory, raising=False)
    dev = torch.device("cuda", 0)
    over = ["Model.name=small_mocov3_pretrain", "Global.epochs=1", "Global.output_dir=%s" % tmp_path, "Global.print_batch_step=1",
            "LRScheduler.warmup_epoch=1", "LRScheduler.warmup_start_lr=0.0005", "LRScheduler.learning_rate=0.001",
            "DataLoader.Train.sampler.batch_size=32"]

    def make(accum):
        torch.manual_seed(3)
        cfg = get_config(CFG, over)
        cfg["Global"]["max_train_step"] = None
        cfg["Global"]["accum_steps"] = accum
        cfg["Optimizer"]["grad_clip"] = dict(name="ClipGradByGlobalNorm", clip_norm=0.05)
        return Engine(cfg, dataloader=SyntheticTwoViewLists(32, 1, dev, size=64))
    e = make(2)
    w0 = e.store.master.clone()
    batch = next(iter(e.train_dataloader))
    loss = e.train_one_step(batch)
    torch.cuda.synchronize()
    gc = e.optimizer.grad_control
    assert gc is not None and float(gc.found_inf.item()) == 0.0
    # manual reference: same model state, the two halves by hand
    r = make(1)
    assert torch.equal(r.store.master, w0)
    r.optimizer.clear_grad()
    losses = []
    for idx in range(2):
        sub = [b[idx * 16:(idx + 1) * 16].contiguous() for b in batch]
        out = r.model(sub)
        l = out["loss"] if isinstance(out, dict) else out
        l.backward()
        losses.append(float(l.item()))
    g = r.store.grad.clone() / 2
    norm = g.norm().item()
    assert abs(float(gc.global_norm.item()) - norm) <= 1e-3 * norm + 1e-6, (float(gc.global_norm.item()), norm)
    assert norm > 0.05                                                 # the clip is active in this configuration
    assert abs(float(loss.item()) - sum(losses) / 2) <= 1e-3 * abs(sum(losses) / 2)
    r.optimizer.grad_scale = 0.5
    r.optimizer.step()
    torch.cuda.synchronize()
    assert ((e.store.master - r.store.master).norm() / (r.store.master - w0).norm()).item() < 2e-2
    assert not torch.equal(e.store.master, w0)
