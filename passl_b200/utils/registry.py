"""Name -> class registry with the reference's API (passl_v110/utils/registry.py:25-133)."""
import inspect


class Registry(object):
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, \
            "An object named '{}' was already registered in '{}' registry!".format(name, self._name)
        self._obj_map[name] = obj

    def register(self, obj=None, name=None):
        if obj is None:
            def deco(func_or_class, name=name):
                self._do_register(name or func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(name or obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError("No object named '{}' found in '{}' registry!".format(name, self._name))
        return ret

    def __contains__(self, name):
        return name in self._obj_map


def build_from_config(cfg, registry, default_args=None):
    """Instantiate ``registry[cfg['name']](**rest_of_cfg)`` — same contract and errors as the reference (:88-133)."""
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'name' not in cfg and (default_args is None or 'name' not in default_args):
        raise KeyError(f'`cfg` or `default_args` must contain the key "name", but got {cfg}\n{default_args}')
    if not isinstance(registry, Registry):
        raise TypeError(f'registry must be an Registry object, but got {type(registry)}')
    if not (isinstance(default_args, dict) or default_args is None):
        raise TypeError(f'default_args must be a dict or None, but got {type(default_args)}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    cls_name = args.pop('name')
    if isinstance(cls_name, str):
        obj_cls = registry.get(cls_name)
    elif inspect.isclass(cls_name):
        obj_cls = cls_name
    else:
        raise TypeError(f'name must be a str or valid name, but got {type(cls_name)}')
    return obj_cls(**args)
