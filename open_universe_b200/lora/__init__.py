"""LoRA adapters for Conv1d / ConvTranspose1d / Linear layers (reference ``open_universe/lora``).

Inference surface only: the adapters are parameter containers with the reference's attribute names
and ``state_dict`` keys; ``engine.fold.effective_weight`` merges  W + alpha / rank * A B  at load time,
so an adapted model runs through exactly the kernels of the plain one.  Training the adapters
(autograd through the CUDA path) is out of scope."""
from .lora import LoraConv1d, LoraConvTranspose1d, LoraLinear
from .utils import freeze_parameters_except_lora_and_bias, inject, remove

__all__ = ["LoraConv1d", "LoraConvTranspose1d", "LoraLinear", "inject", "remove",
           "freeze_parameters_except_lora_and_bias"]
