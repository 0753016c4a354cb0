from .moco import MoCo  # noqa: F401
from .simclr import SimCLR  # noqa: F401
