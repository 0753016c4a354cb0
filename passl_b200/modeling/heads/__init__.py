from .contrastive_head import ContrastiveHead  # noqa: F401
