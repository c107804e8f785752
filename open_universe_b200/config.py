"""Minimal Hydra/OmegaConf-compatible config handling for the enhance() drop-in.

The reference builds its models with ``hydra.utils.instantiate(cfg, _recursive_=False)`` on
OmegaConf nodes (reference ``inference_utils/model_loader.py:51-59,112-114``;
``networks/universe/universe.py:90-97``).  Neither package is required here: ``Config`` is
an attribute-dict that supports the subset of DictConfig the hot path uses (``cfg.key``,
``cfg["key"]``, ``cfg.get(key, default)``, iteration), ``resolve`` handles ``${a.b.c}``
interpolations and ``instantiate`` imports ``_target_`` -- mapping the reference's
``open_universe.*`` targets onto this package so that the reference's YAML files and
checkpoints' ``config.yaml`` work unchanged.
"""
import importlib
from pathlib import Path

import yaml

PACKAGE = __name__.rsplit(".", 1)[0]
CONFIG_DIR = Path(__file__).resolve().parent / "configs"

MODEL_CONFIGS = {
    "universepp_16k": "universepp_16k.yaml",
    "universe_original_16k": "universe_original_16k.yaml",
    "universepp_24k": "universepp_24k.yaml",
}


class Config(dict):
    """Attribute-access dict (stand-in for omegaconf.DictConfig on the inference path)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]


def to_config(node):
    if isinstance(node, Config):
        return node
    if isinstance(node, dict):
        return Config({k: to_config(v) for k, v in node.items()})
    if isinstance(node, (list, tuple)):
        return [to_config(v) for v in node]
    return node


def to_container(node):
    if isinstance(node, dict):
        return {k: to_container(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [to_container(v) for v in node]
    return node


def _lookup(root, path):
    cur = root
    for part in path.split("."):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def resolve(node, root=None, defaults=None):
    """Resolve absolute ``${a.b.c}`` interpolations against ``root`` (then ``defaults``)."""
    root = node if root is None else root
    if isinstance(node, dict):
        return {k: resolve(v, root, defaults) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v, root, defaults) for v in node]
    if isinstance(node, str) and node.startswith("${") and node.endswith("}"):
        path = node[2:-1]
        for src in (root, defaults or {}):
            try:
                return resolve(_lookup(src, path), root, defaults)
            except (KeyError, IndexError, TypeError):
                continue
        return None  # training-only references (datamodule / trainer) are not needed here
    return node


def _coerce_numbers(model):
    """PyYAML reads ``5e-4`` (no dot) as a string; OmegaConf reads a float."""
    diff = model.get("diffusion")
    if isinstance(diff, dict):
        for key in ("sigma_min", "sigma_max", "epsilon"):
            if isinstance(diff.get(key), str):
                diff[key] = float(diff[key])
    return model


def load_config(path):
    """Open a reference-style ``config.yaml`` (top level has a ``model:`` section) and return a
    resolved ``Config`` (reference ``model_loader.py:51-59``)."""
    with open(path, "r") as f:
        raw = yaml.safe_load(f)
    raw = resolve(raw, raw)
    if "model" in raw:
        _coerce_numbers(raw["model"])
    return to_config(raw)


def builtin_config(name):
    """One of the three model configurations shipped with the reference."""
    if name not in MODEL_CONFIGS:
        raise KeyError(f"unknown model config {name!r}; known: {sorted(MODEL_CONFIGS)}")
    return load_config(CONFIG_DIR / MODEL_CONFIGS[name])


def _import_target(target):
    module, _, name = target.rpartition(".")
    if module == "open_universe" or module.startswith("open_universe."):
        module = PACKAGE + module[len("open_universe"):]
    return getattr(importlib.import_module(module), name)


def instantiate(config=None, _recursive_=True, _convert_=None, **overrides):
    """``hydra.utils.instantiate`` work-alike: import ``_target_`` and call it with the other keys."""
    if config is None:
        return None
    cfg = dict(config)
    cfg.update(overrides)
    target = cfg.pop("_target_")
    cls = _import_target(target) if isinstance(target, str) else target
    if _recursive_:
        cfg = {k: (instantiate(v) if isinstance(v, dict) and "_target_" in v else v)
               for k, v in cfg.items()}
    else:
        cfg = {k: to_config(v) for k, v in cfg.items()}
    return cls(**cfg)
