from .param_store import ParamStore  # noqa: F401
