from .moco import MoCo  # noqa: F401
from .simclr import SimCLR  # noqa: F401
from .clip_wrapper import CLIP, CLIPHead, CLIPWrapper  # noqa: F401
from .mae_wrapper import MAE, MAE_PRETRAIN  # noqa: F401
