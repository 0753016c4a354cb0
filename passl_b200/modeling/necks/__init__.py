from .base_neck import LinearNeck, NonLinearNeckV1, NonLinearNeckfc3  # noqa: F401
