"""``load_model`` -- same contract as the reference's ``inference_utils/model_loader.py:62-137``.

A local checkpoint path is paired with ``./config.yaml`` or ``../.hydra/config.yaml``
(model_loader.py:33-48); anything else is treated as a Hugging Face id ``repo[:revision]`` holding
``weights.ckpt`` + ``config.yaml`` (:84-102).  The checkpoint is a ``torch.load`` dict with
``state_dict`` and optionally ``ema`` (torch_ema state); EMA weights are preferred (:119-130).
The model is returned in ``eval()`` mode on ``device``.

Differences from upstream, all on the safe side:
* training-only ``loss_*`` entries of the checkpoint are ignored (the discriminators / MDN heads are
  not instantiated here);
* ``torch.load`` is asked for ``weights_only`` tensors; a checkpoint that needs full unpickling
  (arbitrary code execution) is only accepted from a LOCAL path, or with ``OU_ALLOW_PICKLE=1`` for
  hub downloads;
* a checkpoint WITHOUT an ``"ema"`` entry: upstream calls ``ema.store(...)`` (model_loader.py:124-127)
  and the following ``eval()`` then copies the EMA shadow -- still the constructor's random
  initialisation -- over the weights it has just loaded.  Here the shadow is set to the loaded
  weights instead, so that ``eval()`` / ``train()`` round-trip them.
"""
import os
import pickle
from pathlib import Path

import torch

from ..config import instantiate, load_config

supported_models = ["universe"]

TRAINING_ONLY_PREFIXES = ("loss_",)


def ckpt_to_config_path(ckpt_path):
    ckpt_path = Path(ckpt_path)
    for cand in (ckpt_path.parent / "config.yaml", ckpt_path.parents[1] / ".hydra/config.yaml"):
        if cand.exists():
            return cand
    raise ValueError(f"Could not find the configuration file for model {ckpt_path}.")


def open_update_config(path):
    return load_config(path)


def _torch_load(path, device, trusted):
    try:
        return torch.load(path, map_location=device, weights_only=True)
    except pickle.UnpicklingError:
        if not (trusted or os.environ.get("OU_ALLOW_PICKLE", "0") == "1"):
            raise
        return torch.load(path, map_location=device, weights_only=False)


def _inference_state_dict(state_dict):
    return {k: v for k, v in state_dict.items() if not k.startswith(TRAINING_ONLY_PREFIXES)}


def load_model(ckpt_path, device=None, strict=True, return_config=False, hf_token=None):
    local = Path(ckpt_path).exists()
    if not local:
        try:
            from huggingface_hub import hf_hub_download
            ckpt_path = str(ckpt_path)
            repo_id, _, revision = ckpt_path.partition(":")
            revision = revision or None
            ckpt_path = hf_hub_download(repo_id=repo_id, filename="weights.ckpt", revision=revision,
                                        token=hf_token)
            config_path = hf_hub_download(repo_id=repo_id, filename="config.yaml",
                                          revision=revision, token=hf_token)
        except Exception as e:
            print(f"{ckpt_path} is not a local file and download from HF hub failed.")
            raise e
    else:
        ckpt_path = Path(ckpt_path)
        config_path = ckpt_to_config_path(ckpt_path)

    config = open_update_config(config_path)
    model = instantiate(config.model, _recursive_=False)
    model = model.to(device)
    data = _torch_load(ckpt_path, device, trusted=local)
    state_dict = _inference_state_dict(data["state_dict"])

    ema = getattr(model, "ema", None)
    if ema is not None and "ema" in data:
        ema.load_state_dict(data["ema"])
        model.load_state_dict(state_dict, strict=False)  # EMA weights are what inference uses
    elif ema is not None:
        model.load_state_dict(state_dict, strict=strict)
        ema.shadow_params = [p.clone().detach() for p in model.model_parameters()]
    else:
        model.load_state_dict(state_dict, strict=strict)
    model.eval()
    if return_config:
        return model, config
    return model
