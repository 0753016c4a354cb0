"""YAML config files -> attribute-style nested mappings, with `-o a.b.0.c=value` overrides from the command line.

Behaviour follows the reference's loader (passl/utils/config.py:24-173) because the YAML files are shared with it: string scalars
that are Python literals become those literals ("(0.9, 0.999)" -> tuple, "1e-8" -> float, "None" -> None; "1.0/255.0" stays a
string), override paths may index into lists, an override may create keys that the file does not have (it says so), and override
values are parsed the same way as file values (plus YAML's own `null` / `true` / `false`)."""
import argparse
import ast
import copy
import os

import yaml


class AttrDict(dict):
    """dict whose keys are also attributes (cfg.model.backbone.depth)."""

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return AttrDict((copy.deepcopy(k, memo), copy.deepcopy(v, memo)) for k, v in self.items())


def _literal(text):
    """A string that spells a Python literal -> the literal; anything else unchanged."""
    if not isinstance(text, str):
        return text
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError, TypeError, MemoryError, RecursionError):
        return text


def _attrify(node):
    """Recursively: mappings -> AttrDict, string leaves -> literals.  Lists keep their identity (their mapping items are converted)."""
    if isinstance(node, dict):
        return AttrDict((k, _attrify(v)) for k, v in node.items())
    if isinstance(node, list):
        return [_attrify(v) if isinstance(v, (dict, list)) else v for v in node]
    return _literal(node)


def _override_value(text):
    value = _literal(text)
    if isinstance(value, str):                       # not a Python literal: let YAML read null / true / plain words
        try:
            parsed = yaml.safe_load(value)
        except yaml.YAMLError:
            return value
        return value if isinstance(parsed, (dict, list)) else parsed
    return value


def _set_by_path(root, path, value, option):
    node = root
    for depth, part in enumerate(path):
        last = depth == len(path) - 1
        if isinstance(node, list):
            index = int(part)
            if not -len(node) <= index < len(node):
                raise IndexError("override %r: index %d is outside a list of %d items" % (option, index, len(node)))
            if last:
                node[index] = value
            else:
                node = node[index]
        elif isinstance(node, dict):
            if last:
                if part not in node:
                    print("override %r adds the new key %r" % (option, part))
                node[part] = value
            else:
                if part not in node:
                    node[part] = AttrDict()
                node = node[part]
        else:
            raise TypeError("override %r: %r is a %s, cannot descend into it" % (option, ".".join(path[:depth]), type(node).__name__))


def override_config(config, options=None):
    for option in options or []:
        if not isinstance(option, str) or option.count("=") < 1:
            raise ValueError("override %r must look like key.path=value" % (option,))
        key, _, text = option.partition("=")
        if not key:
            raise ValueError("override %r has an empty key" % (option,))
        _set_by_path(config, key.split("."), _override_value(text), option)
    return config


def parse_config(path):
    with open(path, "r") as f:
        return _attrify(yaml.safe_load(f) or {})


def get_config(fname, overrides=None, show=False):
    if not os.path.exists(fname):
        raise FileNotFoundError("config file %s does not exist" % fname)
    return override_config(parse_config(fname), overrides)


def parse_args(argv=None):
    """The train entry's flags: -c / -o / -p of the v2.5 CLI (passl/utils/config.py:151-173), --resume / --load of the v110 one
    (passl_v110/utils/options.py:35-49; --evaluate-only / --export are outside the training path)."""
    ap = argparse.ArgumentParser("passl_b200 train script")
    ap.add_argument("-c", "--config", "--config-file", dest="config", type=str, default="configs/config.yaml", help="config file path")
    ap.add_argument("-o", "--override", action="append", default=[], help="key.path=value, repeatable")
    ap.add_argument("-p", "--profiler_options", type=str, default=None, help='step window, e.g. "batch_range=[10, 20]; profile_path=run.txt"')
    ap.add_argument("--resume", type=str, default=None, help="checkpoint to continue from (weights, schedule position, iteration)")
    ap.add_argument("--load", type=str, default=None, help="weights to start from (.pdparams / epoch_N.pd / this package's iter_N.pd)")
    return ap.parse_args(argv)
