"""Import shim that lets the UNMODIFIED reference hot-path modules run in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (run by hand in the
build container, where ``/root/reference`` is mounted) to produce the committed golden
vectors.  Nothing in the product, in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports
this file; ``/root/reference`` does not exist on the GPU box.

The reference (``/root/reference/open_universe``) imports four packages that are not
installed here -- pytorch_lightning, hydra, omegaconf, torch_ema -- and its top-level
``__init__`` pulls in metric packages (onnxruntime, pesq, ...).  Following SURVEY.md
Appendix A we register minimal stand-ins in ``sys.modules`` *before* importing and expose
``open_universe`` as a bare namespace whose ``__path__`` points at the reference tree, so
that ``open_universe.networks.universe.*`` imports the reference's own files unchanged.
"""
import importlib
import itertools
import sys
import types
from pathlib import Path

import torch
import yaml

REF_ROOT = Path("/root/reference")


class AttrDict(dict):
    """dict with attribute access and ``.get`` -- enough of DictConfig for the hot path."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(x):
    if isinstance(x, dict):
        return AttrDict({k: to_attr(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return [to_attr(v) for v in x]
    return x


def _instantiate(config=None, _recursive_=True, _convert_=None, **overrides):
    cfg = dict(config)
    cfg.update(overrides)
    target = cfg.pop("_target_")
    mod, _, name = target.rpartition(".")
    cls = getattr(importlib.import_module(mod), name)
    if _recursive_:
        cfg = {
            k: (_instantiate(v) if isinstance(v, dict) and "_target_" in v else v)
            for k, v in cfg.items()
        }
    return cls(**cfg)


class _EMA:
    """torch_ema.ExponentialMovingAverage look-alike (store/copy_to/restore only)."""

    def __init__(self, parameters, decay):
        self.decay = decay
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.collected_params = None

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters]

    def copy_to(self, parameters):
        for s, p in zip(self.shadow_params, parameters):
            p.data.copy_(s.data)

    def restore(self, parameters):
        for c, p in zip(self.collected_params, parameters):
            p.data.copy_(c.data)
        self.collected_params = None

    def update(self, parameters):
        pass

    def to(self, *a, **k):
        self.shadow_params = [p.to(*a, **k) for p in self.shadow_params]

    def state_dict(self):
        return {"decay": self.decay, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params, "num_updates": 0}

    def load_state_dict(self, sd):
        self.shadow_params = [p.clone() for p in sd["shadow_params"]]


def install():
    if "open_universe" in sys.modules:
        return
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    class LightningDataModule:
        pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    sys.modules["pytorch_lightning"] = pl

    hydra = types.ModuleType("hydra")
    hutils = types.ModuleType("hydra.utils")
    hutils.instantiate = _instantiate
    hutils.to_absolute_path = lambda p: p
    hydra.utils = hutils
    sys.modules["hydra"] = hydra
    sys.modules["hydra.utils"] = hutils

    oc = types.ModuleType("omegaconf")

    class OmegaConf:
        @staticmethod
        def create(x):
            return to_attr(x)

        @staticmethod
        def to_container(x, resolve=True):
            return x

    oc.OmegaConf = OmegaConf
    oc.DictConfig = AttrDict
    sys.modules["omegaconf"] = oc

    te = types.ModuleType("torch_ema")
    te.ExponentialMovingAverage = _EMA
    sys.modules["torch_ema"] = te

    pkg = types.ModuleType("open_universe")
    pkg.__path__ = [str(REF_ROOT / "open_universe")]
    sys.modules["open_universe"] = pkg


def _resolve(node, root):
    if isinstance(node, dict):
        return {k: _resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root) for v in node]
    if isinstance(node, str) and node.startswith("${") and node.endswith("}"):
        cur = root
        for part in node[2:-1].split("."):
            cur = cur[part]
        return _resolve(cur, root)
    return node


def load_model_config(name):
    """Return the resolved ``model`` section of ``config/model/<name>.yaml``."""
    with open(REF_ROOT / "config" / "model" / f"{name}.yaml") as f:
        model = yaml.safe_load(f)
    root = {
        "model": model,
        "datamodule": {"datasets": {"vb-train-16k": {"audio_len": 2.0},
                                     "distorted-speech": {"speech_len": 2.0}}},
        "trainer": {"max_steps": 600000},
    }
    model = _resolve(model, root)
    model["validation"]["enh_losses"] = {}
    model["diffusion"]["sigma_min"] = float(model["diffusion"]["sigma_min"])
    model["diffusion"]["sigma_max"] = float(model["diffusion"]["sigma_max"])
    return model


def build_reference_model(name):
    """Instantiate the reference model class for a model YAML (random init, eval mode)."""
    install()
    cfg = to_attr(load_model_config(name))
    model = _instantiate(cfg, _recursive_=False)
    return model, cfg


def set_injected_noise(noise_list):
    """Replace the reference's module-level ``randn`` (universe.py:39-41) by a pop-from-list."""
    install()
    u = importlib.import_module("open_universe.networks.universe.universe")
    it = iter(noise_list)

    def randn(x, sigma, rng=None):
        n = next(it)
        assert n.shape == x.shape, (n.shape, x.shape)
        return n.to(x) * sigma[:, None, None]

    u.randn = randn
