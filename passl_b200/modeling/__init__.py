from .registry import (MODELS, BACKBONES, NECKS, HEADS, build_model, build_backbone, build_neck, build_head)  # noqa: F401
from . import backbones, necks, heads, architectures  # noqa: F401  (populate the registries)
