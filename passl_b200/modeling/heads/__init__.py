from .contrastive_head import ContrastiveHead  # noqa: F401
from .simclr_contrastive_head import SimCLRContrastiveHead  # noqa: F401
