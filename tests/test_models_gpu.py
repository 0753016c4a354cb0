"""Model-level smoke + invariants on the GPU: MoCo v3 (ViT + MLP heads + cosine EMA), MAE at ViT-B/16 shape, Trainer surface."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_mocov3_small_step():
    from passl_b200.models.mocov3 import MoCoV3Pretrain, MoCoV3ViT
    from passl_b200.optimizer import AdamW

    def enc():
        return MoCoV3ViT(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, qkv_bias=True, epsilon=1e-6)
    torch.manual_seed(0)
    m = MoCoV3Pretrain(enc, dim=128, mlp_dim=256, T=0.2, max_steps=10).cuda()
    st, sk = m.build_param_stores()
    opt = AdamW(st, lr=1e-3, weight_decay=0.1)
    k0 = sk.master.clone()
    losses = []
    for it in range(3):
        x1 = torch.randn(16, 3, 64, 64, device="cuda")
        x2 = x1 + 0.1 * torch.randn_like(x1)
        opt.clear_grad()
        loss = m([x1, x2])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)), losses
    # CE*2T over 16 keys: loss <= 2 * 2T * ln(16) at random init, > 0
    assert 0 < losses[0] < 2 * 2 * 0.2 * np.log(16) + 0.5
    assert not torch.equal(sk.master, k0)                       # momentum encoder moved
    assert st.grad.abs().sum().item() > 0
    # the momentum encoder never receives gradients
    assert all(p.grad is None for p in m.momentum_encoder.parameters())


def test_mae_vit_base_shape_step():
    """BASELINE config C4 shapes at a small batch: ViT-B/16 encoder on 50 tokens, 512-d decoder on 197 tokens (d=32 heads)."""
    from passl_b200.core import ParamStore
    from passl_b200.models import build_model
    from passl_b200.optimizer import AdamW
    torch.manual_seed(0)
    m = build_model(dict(name="mae_vit_base_patch16", norm_pix_loss=True)).cuda()
    st = ParamStore(m)
    opt = AdamW(st, lr=1.5e-4, beta2=0.95, weight_decay=0.05)
    imgs = torch.randn(8, 3, 224, 224, device="cuda")
    l0 = None
    for it in range(3):
        opt.clear_grad()
        loss, pred, mask = m(imgs, mask_ratio=0.75)
        loss.backward()
        opt.step()
        assert np.isfinite(loss.item())
        l0 = l0 or loss.item()
    assert pred.shape == (8, 196, 768) and mask.shape == (8, 196)
    assert mask.sum().item() == 8 * 147                          # B * (196 - int(196*0.25))
    assert 0.5 < l0 < 3.0                                        # norm-pix MSE of an untrained model is ~1


def test_trainer_surface_moco(tmp_path):
    from passl_b200.engine.trainer import Trainer
    from passl_b200.utils.config import get_config
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"),
                     ["model.K=1024", "dataloader.train.sampler.batch_size=16", "total_iters=4", "log_config.interval=2"])
    from passl_b200.engine import trainer as T
    tr = Trainer(cfg, dataloader=T.SyntheticTwoViews(16, 4, torch.device("cuda"), size=64))
    out = tr.train()
    assert np.isfinite(float(out["loss"])) and "acc1" in out and "acc5" in out
    tr.model.flush_queue()
    assert int(tr.model.queue_ptr.item()) == (4 * 16) % 1024


def test_trainer_checkpoint_resume(tmp_path):
    """save -> fresh Trainer -> resume reproduces parameters, optimizer state, queue pointer and the next step's loss."""
    from passl_b200.engine import trainer as T
    from passl_b200.utils.config import get_config
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    over = ["model.K=1024", "dataloader.train.sampler.batch_size=16", "total_iters=2", "log_config.interval=100",
            "output_dir=%s" % tmp_path]
    cfg = get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), over)
    data = T.SyntheticTwoViews(16, 2, torch.device("cuda"), size=64)
    tr = T.Trainer(cfg, dataloader=data)
    tr.train()
    path = tr.save()
    tr.model.flush_queue()
    ptr_saved = int(tr.model.queue_ptr.item())
    # one more step on the original
    tr.dataloader = T.SyntheticTwoViews(16, 3, torch.device("cuda"), size=64)
    out_a = tr.train()
    cfg2 = get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), over)
    tr2 = T.Trainer(cfg2, dataloader=T.SyntheticTwoViews(16, 3, torch.device("cuda"), size=64))
    tr2.resume(path)
    assert tr2.current_iter == 2 and int(tr2.model.queue_ptr.item()) == ptr_saved
    out_b = tr2.train()
    np.testing.assert_allclose(float(out_b["loss"]), float(out_a["loss"]), rtol=2e-3)
    torch.testing.assert_close(tr2.store.master, tr.store.master, rtol=1e-3, atol=1e-4)


def test_trainer_surface_clip():
    """configs/clip (reference keys) -> Trainer with the synthetic image-text loader: a few AdamW steps on a shrunken CLIP."""
    from passl_b200.engine.trainer import Trainer
    from passl_b200.utils.config import get_config
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = "model.architecture."
    cfg = get_config(os.path.join(root, "configs/clip/vit-b-16.yaml"),
                     [a + "vision_layers=2", a + "vision_width=128", a + "image_resolution=64", a + "embed_dim=64",
                      a + "transformer_width=128", a + "transformer_heads=2", a + "transformer_layers=2", a + "context_length=16",
                      a + "vocab_size=1000", "dataloader.train.sampler.batch_size=16", "total_iters=3", "log_config.interval=100"])
    tr = Trainer(cfg)
    out = tr.train()
    assert np.isfinite(float(out["loss"])) and np.isfinite(float(out["img_loss"])) and np.isfinite(float(out["text_loss"]))
    assert 0 < float(out["loss"]) < 2 * np.log(16) + 2.0
    assert abs(tr.model.model.logit_scale.item()) <= 4.6 + 1e-6
