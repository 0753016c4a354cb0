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
