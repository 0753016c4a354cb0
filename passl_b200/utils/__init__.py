from .registry import Registry, build_from_config  # noqa: F401
