"""Name -> class tables behind the config-driven construction (`backbone: {name: ResNet, depth: 50}`).

Same surface as the reference's registries (passl_v110/utils/registry.py:25-133) because the YAML files and the `@X.register()`
decorators are shared vocabulary: `Registry(name)`, `.register()` as decorator or call, `.get(name)`, and
`build_from_config(cfg, registry, default_args)` = look the class up by `cfg['name']` and call it with the remaining keys."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._table = {}

    def __repr__(self):
        return "Registry(%r, %d entries)" % (self._name, len(self._table))

    def __contains__(self, key):
        return key in self._table

    def names(self):
        return sorted(self._table)

    def _add(self, obj, key=None):
        key = key or obj.__name__
        if key in self._table:
            raise AssertionError("An object named '%s' was already registered in '%s' registry!" % (key, self._name))
        self._table[key] = obj
        return obj

    def register(self, obj=None, name=None):
        """`@R.register()`, `@R.register(name='x')` or `R.register(cls)`."""
        if obj is not None:
            self._add(obj, name)
            return None
        return lambda target: self._add(target, name)

    def get(self, name):
        try:
            return self._table[name]
        except KeyError:
            raise KeyError("No object named '%s' found in '%s' registry! (known: %s)" % (name, self._name, ", ".join(self.names()))) from None


def build_from_config(cfg, registry, default_args=None):
    """cfg: mapping with a 'name' (a registered name, or a class) plus constructor keywords; default_args fill missing keywords."""
    if not isinstance(cfg, dict):
        raise TypeError("cfg must be a dict, but got %s" % type(cfg))
    if not isinstance(registry, Registry):
        raise TypeError("registry must be an Registry object, but got %s" % type(registry))
    if default_args is not None and not isinstance(default_args, dict):
        raise TypeError("default_args must be a dict or None, but got %s" % type(default_args))
    kwargs = dict(default_args or {})
    kwargs.update(cfg)
    if "name" not in kwargs:
        raise KeyError('`cfg` or `default_args` must contain the key "name", but got %s\n%s' % (cfg, default_args))
    target = kwargs.pop("name")
    if isinstance(target, str):
        target = registry.get(target)
    elif not isinstance(target, type):
        raise TypeError("name must be a str or valid name, but got %s" % type(target))
    return target(**kwargs)
