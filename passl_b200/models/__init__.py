"""v2.5 model registry surface: `build_model(config)` = `getattr(passl_b200.models, name)(**config)`
(passl/models/__init__.py:37-44)."""
import copy

from .vision_transformer import VisionTransformer, ViT_base_patch16_224  # noqa: F401
from .mae import MaskedAutoencoderViT, mae_vit_base_patch16  # noqa: F401
from .mocov3 import MoCoV3Pretrain, MoCoV3ViT, mocov3_vit_base_pretrain  # noqa: F401
from .clip import CLIP, CLIPHead, CLIPWrapper, CLIPVisionTransformer, CLIPTextTransformer  # noqa: F401
from ..modeling.backbones.resnet import ResNet  # noqa: F401


def resnet50(**kwargs):
    """passl/models/resnet.py:95-214 factory (feature extractor; zero-init last BN gamma optional)."""
    return ResNet(depth=50, **kwargs)


def build_model(config):
    config = copy.deepcopy(dict(config))
    model_type = config.pop("name")
    import sys
    mod = sys.modules[__name__]
    model = getattr(mod, model_type)(**config)
    return model
