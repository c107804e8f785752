"""Adapter modules (reference lora/lora.py:23-237): same constructor checks, parameter names, shapes
and initialisation; the merged weight is consumed by ``engine.fold``, there is no ``forward``."""
from typing import Optional

import torch


class _Adapter(torch.nn.Module):
    inner_name = "conv"
    base_type = torch.nn.Conv1d

    def __init__(self, module: torch.nn.Module, rank: int, alpha: Optional[float] = None):
        super().__init__()
        if not isinstance(module, self.base_type):
            raise ValueError(f"module should be an instance of {self.base_type.__name__}")
        if getattr(module, "padding_mode", "zeros") != "zeros":
            raise ValueError("LoRA only supports padding_mode='zeros'")
        if hasattr(module, "weight_g"):
            raise ValueError("remove weight norm before injecting LoRA adapters (lora.py:159)")
        self.rank = rank
        self.alpha = alpha if alpha is not None else rank
        setattr(self, self.inner_name, module)
        w = module.weight
        if w.shape[1] < rank or w.shape[0] < rank:
            raise ValueError("The rank should be smaller than the input and output size")
        self._init_factors(w)

    def lora_inner(self):
        return getattr(self, self.inner_name)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("LoRA adapters are merged into the packed weights of the CUDA path; "
                                  "call the owning network, not the layer")


class LoraConv1d(_Adapter):
    """lora.py:23-97: W (out, in, k) + alpha/rank * (A (out, r) @ B (r, in*k)).view_as(W); A = 0, B ~ N(0,1)."""

    def _init_factors(self, w):
        self.lora_weight_a = torch.nn.Parameter(w.new_zeros(w.shape[0], self.rank))
        self.lora_weight_b = torch.nn.Parameter(w.new_zeros(self.rank, w.shape[1] * w.shape[2]).normal_())

    def lora_merged_weight(self):
        w = self.lora_inner().weight
        lora_w = (self.lora_weight_a @ self.lora_weight_b).view(w.shape)
        return w + (self.alpha / self.rank) * lora_w

    _get_weights = lora_merged_weight

    def un_lora(self):
        m = self.lora_inner()
        conv = type(m)(m.in_channels, m.out_channels, m.kernel_size, stride=m.stride, padding=m.padding,
                       dilation=m.dilation, groups=m.groups, bias=m.bias is not None,
                       device=m.weight.device, dtype=m.weight.dtype,
                       **({"output_padding": m.output_padding} if hasattr(m, "output_padding") and
                          isinstance(m, torch.nn.ConvTranspose1d) else {}))
        with torch.no_grad():
            conv.weight.copy_(self.lora_merged_weight())
            if m.bias is not None:
                conv.bias.copy_(m.bias)
        return conv


class LoraConvTranspose1d(LoraConv1d):
    """lora.py:100-178 (weight (in, out, k): the factor shapes follow dims 0 and 1 x 2 as upstream)."""
    base_type = torch.nn.ConvTranspose1d


class LoraLinear(_Adapter):
    """lora.py:181-237: W (out, in) + alpha/rank * A (out, r) @ B (r, in); A ~ N(0,1), B = 0."""
    inner_name = "linear"
    base_type = torch.nn.Linear

    def _init_factors(self, w):
        self.lora_linear_a = torch.nn.Parameter(w.new_zeros(w.shape[0], self.rank).normal_())
        self.lora_linear_b = torch.nn.Parameter(w.new_zeros(self.rank, w.shape[1]))

    def lora_merged_weight(self):
        return self.lora_inner().weight + (self.alpha / self.rank) * (self.lora_linear_a @ self.lora_linear_b)

    _get_weights = lora_merged_weight

    def un_lora(self):
        m = self.lora_inner()
        linear = torch.nn.Linear(m.in_features, m.out_features, bias=m.bias is not None,
                                 device=m.weight.device, dtype=m.weight.dtype)
        with torch.no_grad():
            linear.weight.copy_(self.lora_merged_weight())
            if m.bias is not None:
                linear.bias.copy_(m.bias)
        return linear
