from .moco import MoCo  # noqa: F401
