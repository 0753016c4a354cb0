"""Architecture parity on the CPU (no kernels run): the name-based registries build the reference's hot-path models from the same
config keys, and the parameter counts equal the published / derivable counts of the reference architectures."""
import torch


def _count(m):
    return sum(p.numel() for p in m.parameters())


def test_registry_names():
    from passl_b200.modeling import BACKBONES, HEADS, MODELS, NECKS
    for n in ("MoCo", "SimCLR", "CLIPWrapper"):
        assert n in MODELS, n
    for n in ("ResNet", "ResNetsimclr", "CLIP"):
        assert n in BACKBONES, n
    for n in ("LinearNeck", "NonLinearNeckV1", "NonLinearNeckfc3"):
        assert n in NECKS, n
    for n in ("ContrastiveHead", "SimCLRContrastiveHead", "CLIPHead"):
        assert n in HEADS, n
    import passl_b200.models as M
    for n in ("resnet50", "ViT_base_patch16_224", "mae_vit_base_patch16", "mocov3_vit_base_pretrain", "build_model"):
        assert hasattr(M, n), n


def test_resnet50_parameter_count():
    """torchvision / paddle.vision ResNet-50 without the fc layer: 23,508,032 parameters (resnetimagenet.py:93-246); the stem
    weight is stored padded [64, 152] here (147 real columns)."""
    from passl_b200.modeling import build_backbone
    m = build_backbone(dict(name="ResNet", depth=50))
    assert _count(m) - 64 * (152 - 147) == 23508032
    n_bn = sum(1 for mod in m.modules() if mod.__class__.__name__ == "BatchNormState")
    assert n_bn == 53                                             # one per convolution


def test_moco_v2_config_builds():
    import os
    from passl_b200.modeling import build_model
    from passl_b200.utils.config import get_config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = get_config(os.path.join(root, "configs/moco/moco_v2_r50.yaml"), ["model.K=4096"])
    m = build_model(dict(cfg.model))
    neck = _count(m.encoder_q[1])
    assert neck == 2048 * 2048 + 2048 + 2048 * 128 + 128           # NonLinearNeckV1 2048 -> 2048 -> 128 (base_neck.py:67-94)
    assert m.queue.shape == (4096, 128) and m.queue_ptr.dtype == torch.int64
    assert all(not p.requires_grad for p in m.encoder_k.parameters())
    assert all(mod.use_global_stats for mod in m.encoder_k.modules() if mod.__class__.__name__ == "BatchNormState")


def test_vit_b16_and_mae_parameter_counts():
    import passl_b200.models as M
    vit = M.ViT_base_patch16_224()
    # 12 blocks x 7,087,872 + patch embed 590,592 + cls 768 + pos 151,296 + final norm 1,536 = 85,798,656 (timm vit_base_patch16_224, no head)
    assert _count(vit) == 85798656
    mae = M.mae_vit_base_patch16()
    # encoder 85,798,656 + decoder (embed 393,728; mask token 512; pos 100,864; 8 blocks x 3,152,384; norm 1,024; pred 393,984)
    assert _count(mae) == 85798656 + 393728 + 512 + 197 * 512 + 8 * 3152384 + 1024 + 393984


def test_clip_vit_b16_parameter_count():
    """OpenAI CLIP ViT-B/16: 149,620,737 parameters (vision 86,192,640 incl. proj, text 63,428,096 incl. projection, logit_scale 1)."""
    from passl_b200.modeling import build_model
    arch = dict(name="CLIP", embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=16,
                pre_norm=True, proj=True, patch_bias=False, context_length=77, vocab_size=49408, transformer_width=512,
                transformer_heads=8, transformer_layers=12, qkv_bias=True)
    m = build_model(dict(name="CLIPWrapper", architecture=arch, head=dict(name="CLIPHead")))
    assert _count(m) == 149620737


def test_conv_init_rules_of_the_two_backbones():
    """MoCo's ResNet: kaiming-normal fan_out (backbones/resnet.py:75-80); ResNetsimclr: XavierNormal(fan_out=0) = sqrt(2 / fan_in)
    on every conv, init_parameters() left commented out (resnetcifar.py:137-141, resnetsimclr.py:62)."""
    import math
    import torch
    from passl_b200.modeling.backbones.resnet import ResNet, ResNetsimclr
    torch.manual_seed(0)
    a, b = ResNet(depth=50), ResNetsimclr(depth=50)
    for net, rule in ((a, lambda m: 2.0 / (m.cout * m.k * m.k)), (b, lambda m: 2.0 / (m.cin * m.k * m.k))):
        for name in ("blocks.0.conv1", "blocks.3.conv2", "blocks.15.conv3", "blocks.3.downsample"):
            m = net.get_submodule(name)
            assert abs(m.weight.std().item() / math.sqrt(rule(m)) - 1) < 0.05, (type(net).__name__, name)
    assert abs(a.stem.weight[:, :147].std().item() / math.sqrt(2.0 / (64 * 49)) - 1) < 0.05
    assert abs(b.stem.weight[:, :147].std().item() / math.sqrt(2.0 / 147) - 1) < 0.05
    assert b.stem.weight[:, 147:].abs().sum() == 0 and not b.stem.maxpool and a.stem.maxpool


def test_unbuilt_constructor_options_fail_loudly():
    """Dropout / stochastic depth / frozen stages are not built: asking for them raises instead of being ignored."""
    import pytest
    from passl_b200.modeling.backbones.resnet import ResNet
    from passl_b200.models.vision_transformer import VisionTransformer
    VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=1, num_heads=2, drop_rate=0.0, drop_path_rate=0, norm_layer="nn.LayerNorm")
    with pytest.raises(NotImplementedError):
        VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=1, num_heads=2, drop_path_rate=0.1)
    with pytest.raises(TypeError):
        VisionTransformer(img_size=32, patch_size=8, embed_dim=64, depth=1, num_heads=2, no_such_option=1)
    with pytest.raises(NotImplementedError):
        ResNet(depth=50, frozen_stages=1)


def test_mocov3_vit_initialisation_rules():
    """mocov3.py:43-61: q / k / v as three Xavier-uniform matrices, cls_token ~ N(0, 1e-6), patch projection uniform with fan
    3 p^2 + D; the MLP heads keep paddle's Linear default (Xavier-uniform, no bias); the momentum encoder starts as a copy."""
    import functools
    import math
    import torch
    from passl_b200.models.mocov3 import MoCoV3Pretrain, MoCoV3ViT
    torch.manual_seed(0)
    D = 256
    m = MoCoV3Pretrain(functools.partial(MoCoV3ViT, img_size=64, patch_size=16, embed_dim=D, depth=2, num_heads=4, qkv_bias=True,
                                         stop_grad_conv1=True), dim=64, mlp_dim=512)
    v = m.base_encoder.vit
    for w, bound in ((v.blocks[1].qkv.weight, math.sqrt(6 / (2 * D))), (v.patch_embed.proj.weight, math.sqrt(6 / (3 * 16 * 16 + D))),
                     (v.blocks[0].fc1.weight, math.sqrt(6 / (D + 4 * D))), (m.predictor.fcs[0].weight, math.sqrt(6 / (64 + 512))),
                     (m.base_encoder.head.fcs[2].weight, math.sqrt(6 / (512 + 64)))):
        assert 0.97 * bound < w.abs().max().item() <= bound
        assert abs(w.std().item() / (bound / math.sqrt(3)) - 1) < 0.05                   # uniform(-b, b): std = b / sqrt(3)
    assert v.cls_token.std().item() < 3e-6 and v.blocks[0].qkv.bias.abs().sum() == 0 and v.patch_embed.proj.bias.abs().sum() == 0
    assert not v.patch_embed.proj.weight.requires_grad and not v.pos_embed.requires_grad
    for a, b in zip(m.base_encoder.parameters(), m.momentum_encoder.parameters()):
        assert torch.equal(a, b) and not b.requires_grad
