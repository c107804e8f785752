"""Shared helpers for the tests: golden loading, deterministic weights, oracle construction."""
import json
from pathlib import Path

import numpy as np
import torch

from cases import MODELS, WEIGHT_SEED  # noqa: F401  (tests/golden on sys.path via conftest)
from detweights import det_audio, det_noise, det_state_dict, is_constructor_buffer  # noqa: F401

GOLDEN = Path(__file__).resolve().parent / "golden"

OUR_CONFIG = {"upp16k": "universepp_16k", "orig16k": "universe_original_16k",
              "upp24k": "universepp_24k"}


def load_manifest(model):
    return json.loads((GOLDEN / f"{model}_manifest.json").read_text())


def load_golden(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


def golden_buffers(model):
    return {k: torch.from_numpy(v) for k, v in np.load(GOLDEN / f"{model}_buffers.npz").items()}


def model_cfg(model):
    from open_universe_b200.config import builtin_config
    return builtin_config(OUR_CONFIG[model]).model


_sd_cache = {}


def full_state_dict(model):
    """Deterministic weights + the reference's constructor-built buffers (from the golden file)."""
    if model not in _sd_cache:
        man = load_manifest(model)["manifest"]
        sd = det_state_dict(man, WEIGHT_SEED)
        sd.update(golden_buffers(model))
        _sd_cache[model] = sd
    return _sd_cache[model]


def make_oracle(model):
    from oracle.universe_oracle import UniverseOracle
    return UniverseOracle(model_cfg(model), full_state_dict(model))


def rel_rms(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).square().mean().sqrt() / b.square().mean().sqrt().clamp(min=1e-30))


def abs_rms(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).square().mean().sqrt())


def sub(t, stride):
    return t.detach().reshape(-1)[:: int(stride)]
