"""Monkey-patching helpers (reference lora/utils.py:24-120)."""
from typing import List, Optional

import torch

from .lora import LoraConv1d, LoraConvTranspose1d, LoraLinear

lora_classes_map = {
    torch.nn.Conv1d: LoraConv1d,
    torch.nn.ConvTranspose1d: LoraConvTranspose1d,
    torch.nn.Linear: LoraLinear,
}


def get_adapter(module, rank, alpha=None):
    """The adapter for ``module``, or None when the layer type / size is not supported (utils.py:31-44)."""
    cls = lora_classes_map.get(type(module))
    if cls is None:
        return None
    try:
        return cls(module, rank, alpha)
    except ValueError:
        return None


def inject(model: torch.nn.Module, rank: int, alpha: Optional[float] = None):
    """Replace every supported Linear / Conv1d / ConvTranspose1d child by its adapter, recursively."""
    for name, module in model.named_children():
        adapter = get_adapter(module, rank, alpha)
        if adapter is not None:
            setattr(model, name, adapter)
        else:
            inject(module, rank, alpha)


def remove(model: torch.nn.Module):
    """Replace the adapters by plain layers holding the merged weights, recursively."""
    for name, module in model.named_children():
        if isinstance(module, (LoraLinear, LoraConv1d)):
            setattr(model, name, module.un_lora())
        else:
            remove(module)


def freeze_parameters_except_lora_and_bias(module: torch.nn.Module, train_biases: Optional[bool] = True,
                                           train_names: Optional[List[str]] = None):
    """requires_grad bookkeeping of the reference (utils.py:88-120): LoRA factors, optionally biases and
    parameters whose name contains one of ``train_names`` stay trainable."""
    train_names = train_names or []
    for name, p in module.named_parameters():
        p.requires_grad = bool("lora_" in name or any(n in name for n in train_names)
                               or (train_biases and "bias" in name))
