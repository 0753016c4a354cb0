"""YAML config -> nested AttrDict with dotted `-o key=value` overrides: same semantics and error messages as
passl/utils/config.py:24-173 (string values literal_eval'd, list indices in override paths, new keys reported)."""
import argparse
import os
from ast import literal_eval

import yaml


class AttrDict(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        if key in self.__dict__:
            self.__dict__[key] = value
        else:
            self[key] = value

    def __deepcopy__(self, memo):
        import copy
        return AttrDict({copy.deepcopy(k, memo): copy.deepcopy(v, memo) for k, v in self.items()})


def create_attr_dict(yaml_config):
    for key, value in yaml_config.items():
        if type(value) is dict:
            yaml_config[key] = value = AttrDict(value)
        if isinstance(value, str):
            try:
                value = literal_eval(value)
            except BaseException:
                pass
        if isinstance(value, AttrDict):
            create_attr_dict(yaml_config[key])
        else:
            yaml_config[key] = value


def parse_config(cfg_file):
    with open(cfg_file, 'r') as fopen:
        yaml_config = AttrDict(yaml.load(fopen, Loader=yaml.SafeLoader))
    create_attr_dict(yaml_config)
    return yaml_config


def override(dl, ks, v):
    def str2num(v):
        try:
            return eval(v)
        except Exception:
            return v
    assert isinstance(dl, (list, dict)), ("{} should be a list or a dict")
    assert len(ks) > 0, ('lenght of keys should larger than 0')
    if isinstance(dl, list):
        k = str2num(ks[0])
        if len(ks) == 1:
            assert k < len(dl), ('index({}) out of range({})'.format(k, dl))
            dl[k] = str2num(v)
        else:
            override(dl[k], ks[1:], v)
    else:
        if len(ks) == 1:
            if not ks[0] in dl:
                print('A new filed ({}) detected!'.format(ks[0], dl))
            dl[ks[0]] = str2num(v)
        else:
            if ks[0] not in dl:
                dl[ks[0]] = AttrDict()
            override(dl[ks[0]], ks[1:], v)


def override_config(config, options=None):
    if options is not None:
        for opt in options:
            assert isinstance(opt, str), ("option({}) should be a str".format(opt))
            assert "=" in opt, ("option({}) should contain a =to distinguish between key and value".format(opt))
            pair = opt.split('=')
            assert len(pair) == 2, ("there can be only a = in the option")
            key, value = pair
            override(config, key.split('.'), value)
    return config


def get_config(fname, overrides=None, show=False):
    assert os.path.exists(fname), ('config file({}) is not exist'.format(fname))
    config = parse_config(fname)
    override_config(config, overrides)
    return config


def parse_args(argv=None):
    """passl/utils/config.py:151-173: -c config, -o overrides, -p profiler options; --resume / --load of the v110 CLI."""
    parser = argparse.ArgumentParser("passl_b200 train script")
    parser.add_argument('-c', '--config', '--config-file', dest='config', type=str, default='configs/config.yaml', help='config file path')
    parser.add_argument('-o', '--override', action='append', default=[], help='config options to be overridden')
    parser.add_argument('-p', '--profiler_options', type=str, default=None, help='profiler options "key1=value1;key2=value2"')
    # passl_v110/utils/options.py:35-49 (training-path subset; --evaluate-only / --export are outside the hot path)
    parser.add_argument('--resume', type=str, default=None, help='checkpoint to continue from (weights, schedule position, iteration)')
    parser.add_argument('--load', type=str, default=None, help='weights to start from (.pdparams / epoch_N.pd / this package\'s iter_N.pd)')
    return parser.parse_args(argv)
