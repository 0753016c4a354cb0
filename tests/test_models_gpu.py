"""Model-level smoke + invariants on the GPU: MoCo v3 (ViT + MLP heads + cosine EMA), MAE at ViT-B/16 shape, Trainer surface."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ParamStoreNumel(module):
    """flat-buffer size of a module: every tensor padded to 1024 elements"""
    return sum((p.numel() + 1023) // 1024 * 1024 for p in module.parameters())


@pytest.mark.parametrize("literal", [False, True])
def test_mocov3_small_step(literal):
    """literal=True: the reference's CosineEMA(Sequential(base_encoder, predictor)) — keys pass through an averaged predictor copy
    and the source is weighted by the cosine momentum (mocov3.py:133-134,224-225, averaged_model.py:165-186)."""
    from passl_b200.models.mocov3 import MoCoV3Pretrain, MoCoV3ViT
    from passl_b200.optimizer import AdamW

    def enc():
        return MoCoV3ViT(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, qkv_bias=True, epsilon=1e-6,
                         stop_grad_conv1=True)
    torch.manual_seed(0)
    m = MoCoV3Pretrain(enc, dim=128, mlp_dim=256, T=0.2, max_steps=10, reference_ema_quirk=literal).cuda()
    st, sk = m.build_param_stores()
    assert sk.numel == (st.numel if literal else ParamStoreNumel(m.base_encoder))
    opt = AdamW(st, lr=1e-3, weight_decay=0.1)
    k0 = sk.master.clone()
    vit = m.base_encoder.vit
    frozen0 = [t.detach().clone() for t in (vit.pos_embed, vit.patch_embed.proj.weight, vit.patch_embed.proj.bias)]
    w0 = vit.blocks[0].qkv.weight.detach().clone()
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
    if literal:
        # after step t the averaged copy is avg (1 - mu_t) + src mu_t with mu_t = 0.99 (cos(pi t / max_steps) + 1) / 2: with the
        # source weighted that heavily the averaged predictor sits within a few per cent of the live one
        pa, pl = m.momentum_predictor.fcs[0].weight.detach(), m.predictor.fcs[0].weight.detach()
        assert not torch.equal(pa, pl) and (pa - pl).norm() < 0.2 * (pl - m.predictor.fcs[0].weight.detach().mean()).norm()
        assert len([k for k in m.state_dict() if k.startswith("_")]) == 0          # enumeration containers stay out of state_dict
    # frozen tensors (fixed sin-cos table, stop_grad_conv1 patch projection: mocov3.py:63-65,91) see neither step nor weight decay
    for t0, t in zip(frozen0, (vit.pos_embed, vit.patch_embed.proj.weight, vit.patch_embed.proj.bias)):
        assert torch.equal(t0, t.detach())
    assert not torch.equal(w0, vit.blocks[0].qkv.weight.detach())


def test_optimizers_leave_frozen_tensors_alone():
    """A frozen tensor in the flat buffer takes no step and no weight decay under any optimizer (the reference's group builder drops
    stop_gradient parameters, passl/optimizer/__init__.py:88-91); its trainable neighbours follow the plain update rule."""
    from passl_b200.core import ParamStore
    from passl_b200.optimizer import AdamW, LarsMomentumOptimizer, Momentum

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Parameter(torch.randn(40, 50))
            self.frozen = torch.nn.Parameter(torch.randn(3, 700), requires_grad=False)
            self.b = torch.nn.Parameter(torch.randn(64, 33))
            self.frozen_tail = torch.nn.Parameter(torch.randn(1, 5, 8), requires_grad=False)
    for make in (lambda s: Momentum(s, lr=0.1, momentum=0.9, weight_decay=0.01),
                 lambda s: LarsMomentumOptimizer(s, lr=0.1, lars_weight_decay=0.01, exclude_from_weight_decay=()),
                 lambda s: AdamW(s, lr=0.1, weight_decay=0.5)):
        torch.manual_seed(5)
        m = M().cuda()
        st = ParamStore(m)
        opt = make(st)
        before = {k: v.detach().clone() for k, v in m.named_parameters()}
        for _ in range(2):
            opt.clear_grad()
            m.a.grad.add_(torch.randn_like(m.a))
            m.b.grad.add_(torch.randn_like(m.b))
            opt.step()
        assert torch.equal(before["frozen"], m.frozen.detach()) and torch.equal(before["frozen_tail"], m.frozen_tail.detach())
        assert not torch.equal(before["a"], m.a.detach()) and not torch.equal(before["b"], m.b.detach())
        assert torch.equal(m.frozen.bf16.float(), m.frozen.detach().bfloat16().float())
    # Momentum rule on the trainable tensors, checked in full: v = mu v + (g + wd p); p -= lr v
    torch.manual_seed(6)
    m = M().cuda()
    st = ParamStore(m)
    opt = Momentum(st, lr=0.1, momentum=0.9, weight_decay=0.01)
    p, v = m.b.detach().clone().double(), torch.zeros_like(m.b, dtype=torch.float64)
    for _ in range(3):
        opt.clear_grad()
        g = torch.randn_like(m.b)
        m.b.grad.add_(g)
        opt.step()
        v = 0.9 * v + (g.double() + 0.01 * p)
        p = p - 0.1 * v
    assert (m.b.detach().double() - p).abs().max().item() < 1e-5


def test_mae_vit_base_shape_step():
    """BASELINE config C4 shapes at a small batch: ViT-B/16 encoder on 50 tokens, 512-d decoder on 197 tokens (d=32 heads)."""
    from passl_b200.core import ParamStore
    from passl_b200.models import build_model
    from passl_b200.optimizer import AdamW
    torch.manual_seed(0)
    m = build_model(dict(name="mae_vit_base_patch16", norm_pix_loss=True)).cuda()
    st = ParamStore(m)
    opt = AdamW(st, lr=1.5e-4, beta2=0.95, weight_decay=0.05, one_dim_no_decay=True)       # MAE recipe (optim_factory.py:21-38)
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
                     ["model.K=1024", "dataloader.train.sampler.batch_size=16", "total_iters=4", "epochs=1", "log_config.interval=2"])
    from passl_b200.engine import trainer as T
    tr = Trainer(cfg, dataloader=T.SyntheticTwoViews(16, 4, torch.device("cuda"), size=64))
    out = tr.train()
    assert np.isfinite(float(out["loss"].detach())) and "acc1" in out and "acc5" in out
    tr.model.flush_queue()
    assert int(tr.model.queue_ptr.item()) == (4 * 16) % 1024


def test_trainer_checkpoint_resume(tmp_path):
    """save -> fresh Trainer -> resume reproduces parameters, optimizer state, queue pointer and the next step's loss."""
    from passl_b200.engine import trainer as T
    from passl_b200.utils.config import get_config
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    over = ["model.K=1024", "dataloader.train.sampler.batch_size=16", "total_iters=2", "epochs=1", "log_config.interval=100",
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


def test_trainer_surface_simclr_recipe():
    """configs/simclr (reference keys): the SimCLR LR recipe (sqrt batch scaling, warm-up from 0, solver/builder.py:46-66) drives LARS
    through the Trainer; the YAML's exclude list matches none of Paddle's generated names, so every tensor keeps its decay."""
    from passl_b200.engine import trainer as T
    from passl_b200.utils.config import get_config
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = get_config(os.path.join(root, "configs/simclr/simclr_r50_IM.yaml"),
                     ["dataloader.train.sampler.batch_size=16", "total_iters=3", "epochs=1", "lr_scheduler.total_images=512",
                      "lr_scheduler.warmup_epochs=1", "log_config.interval=100"])
    tr = T.Trainer(cfg, dataloader=T.SyntheticTwoViews(16, 3, torch.device("cuda"), size=64))
    peak = np.sqrt(16 * 8)                                   # end_lr 1.0 * sqrt(per-GPU batch * 8), engine/trainer.py:161-163
    assert tr.lr_scheduler.warmup_steps == 4 and tr.optimizer.lr == 0.0
    assert int((tr.optimizer.seg_wd == 0).sum()) == 0
    w0 = tr.store.master.clone()
    out = tr.train()
    assert np.isfinite(float(out["loss"].detach())) and torch.isfinite(tr.store.master).all()
    assert abs(tr.optimizer.lr - 0.75 * peak) < 1e-9         # three scheduler steps into a four-step warm-up
    assert not torch.equal(w0, tr.store.master)


def test_trainer_pdparams_weights_roundtrip(tmp_path):
    """Weights written in the reference's container / names / layouts (utils/checkpoint.py) and read back into a fresh Trainer give
    the same parameters (fp32 master and the bf16 mirror the kernels read) and the same loss on the same batch."""
    from passl_b200.engine import trainer as T
    from passl_b200.utils import checkpoint as C
    from passl_b200.utils.config import get_config
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    over = ["model.K=1024", "dataloader.train.sampler.batch_size=16", "total_iters=2", "epochs=1", "log_config.interval=100"]
    tr = T.Trainer(get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), over),
                   dataloader=T.SyntheticTwoViews(16, 2, torch.device("cuda"), size=64))
    tr.train()
    path = tr.save(str(tmp_path / "epoch_1.pdparams"))
    state = C.load_pdparams(path)
    assert state["encoder_q.0.conv1.weight"].shape == (64, 3, 7, 7) and state["queue"].shape == (128, 1024)
    tr2 = T.Trainer(get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), over),
                    dataloader=T.SyntheticTwoViews(16, 2, torch.device("cuda"), size=64))
    tr2.resume(path)
    tr.model.flush_queue()
    for (k, a), (_, b) in zip(tr.model.state_dict().items(), tr2.model.state_dict().items()):
        if k.endswith("stem.weight"):
            a, b = a[:, :147], b[:, :147]
        assert torch.equal(a, b), k
    qa, qb = tr.model.encoder_q[0].blocks[3].conv2.weight, tr2.model.encoder_q[0].blocks[3].conv2.weight
    assert torch.equal(qa.bf16, qb.bf16)                                   # the compute mirror was refreshed
    torch.manual_seed(9)
    x1 = torch.randn(16, 3, 64, 64, device="cuda")
    x2 = x1 + 0.1 * torch.randn_like(x1)
    la = tr.model(x1, x2)["loss"].item()
    lb = tr2.model(x1, x2)["loss"].item()
    assert abs(la - lb) <= 1e-3 * abs(la), (la, lb)
    # the v110 training-checkpoint container (`epoch_N.pd`: pickle with epoch, state_dict, lr_scheduler): position restored too
    path2 = tr.save(str(tmp_path / "epoch_1.pd"), paddle_format=True)
    assert C.is_paddle_pickle(path2) and not C.is_paddle_pickle(tr.save(str(tmp_path / "iter_x.pd")))
    tr3 = T.Trainer(get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), over),
                    dataloader=T.SyntheticTwoViews(16, 2, torch.device("cuda"), size=64))
    tr4 = T.Trainer(get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), over),
                    dataloader=T.SyntheticTwoViews(16, 2, torch.device("cuda"), size=64))
    tr4.load(path2)                                                        # weights only: position / schedule stay as constructed
    assert tr4.current_iter == 0 and torch.equal(tr4.store.master, tr.store.master)
    tr3.resume(path2)
    assert tr3.current_iter == 2 and tr3.lr_scheduler.last_epoch == tr.lr_scheduler.last_epoch and tr3.optimizer.lr == tr.optimizer.lr
    tr.model.flush_queue()
    assert torch.equal(tr3.model.queue, tr.model.queue) and torch.equal(tr3.store.master, tr.store.master)


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
                      a + "vocab_size=1000", "dataloader.train.sampler.batch_size=16", "total_iters=3", "epochs=1", "log_config.interval=100"])
    tr = Trainer(cfg)
    out = tr.train()
    assert np.isfinite(float(out["loss"].detach())) and np.isfinite(float(out["img_loss"])) and np.isfinite(float(out["text_loss"]))
    assert 0 < float(out["loss"].detach()) < 2 * np.log(16) + 2.0
    assert abs(tr.model.model.logit_scale.item()) <= 4.6 + 1e-6
