"""Registries of the v1.1.0 model zoo surface (passl_v110/modeling/{architectures,backbones,necks,heads}/builder.py)."""
from ..utils.registry import Registry, build_from_config

MODELS = Registry("MODEL")
BACKBONES = Registry("BACKBONE")
NECKS = Registry("NECK")
HEADS = Registry("HEAD")


def build_model(cfg):
    return build_from_config(cfg, MODELS)


def build_backbone(cfg):
    return build_from_config(cfg, BACKBONES)


def build_neck(cfg):
    return build_from_config(cfg, NECKS)


def build_head(cfg):
    return build_from_config(cfg, HEADS)
