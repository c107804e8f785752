"""Building blocks of the UNIVERSE networks -- parameter containers for the B200 engine.

Mirrors the public surface of the reference's ``networks/universe/blocks.py`` (same class
names, constructor keywords, attribute names and ``state_dict`` keys, so reference checkpoints
load unchanged) but none of these modules computes anything with ATen: ``forward`` lowers to
fused CUDA kernels through ``open_universe_b200.engine`` (no CPU fallback).

Reference behaviour each class stands for:
  * ``film``              blocks.py:53-59     gamma * x + beta, gamma/beta = halves of a 2C vector
  * ``get_binomial_filter`` blocks.py:62-68   Pascal row scaled to unit RMS
  * ``BinomialAntiAlias`` blocks.py:119-130   depthwise FIR, zero 'same' padding
  * ``PReLU_Conv``        blocks.py:133-227   PReLU -> [low-pass] -> conv/convT -> [low-pass] -> [+bias]
  * ``ConvBlock``         blocks.py:230-412   rate change + conv1/FiLM/conv2/conv3 + residuals
"""
import math
import warnings
from typing import Optional

import torch

from ...engine import runtime


def _weight_norm(module):
    # old-style weight norm keeps the reference's ``weight_g`` / ``weight_v`` checkpoint keys
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(module)


def init_weights(m, mean=0.0, std=0.01):
    if "Conv" in m.__class__.__name__:
        m.weight.data.normal_(mean, std)


def cond_weight_norm(x, use=False):
    """blocks.py:36-41 -- wrap in (old-style) weight norm and re-draw conv weights N(0, 0.01)."""
    if not use:
        return x
    x = _weight_norm(x)
    x.apply(init_weights)
    return x


def remove_weight_norm(model):
    """blocks.py:44-50 -- recursively fold g * v / ||v|| back into ``weight``."""
    for _, child in model.named_children():
        try:
            torch.nn.utils.remove_weight_norm(child)
        except ValueError:
            remove_weight_norm(child)


def film(x, y):
    """FiLM modulation on (B, C, T) fp32 tensors (blocks.py:53-59).  Inside the networks FiLM
    is fused into the conv1 epilogue; this standalone entry point runs the same CUDA math."""
    if y.shape[1] != 2 * x.shape[1]:
        raise ValueError("g should have 2 times more channels than y")
    return runtime.film(x, y)


def get_binomial_filter(kernel_size):
    """Row ``kernel_size-1`` of Pascal's triangle, scaled so that the taps have unit RMS
    (blocks.py:62-68: divided by the RMS of the whole lower-triangular Pascal matrix, then
    re-normalised to unit RMS -- the second step makes the first immaterial)."""
    row = torch.tensor([math.comb(kernel_size - 1, i) for i in range(kernel_size)],
                       dtype=torch.float64)
    row = row / row.square().mean().sqrt()
    return row.to(torch.float32)


class BinomialAntiAlias(torch.nn.Module):
    def __init__(self, kernel_size):
        super().__init__()
        self.register_buffer("weights", get_binomial_filter(kernel_size))

    def forward(self, x):
        return runtime.lowpass(x, self.weights)


class PReLU_Conv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, padding_mode="zeros", device=None, dtype=None,
                 use_transpose=False, prelu_kwargs=None, act_type="prelu",
                 use_weight_norm=False, use_antialiasing=False):
        super().__init__()
        if dilation != 1 or groups != 1 or padding_mode != "zeros":
            raise NotImplementedError("only dense, undilated, zero-padded convs are on the hot path")
        prelu_kwargs = {} if prelu_kwargs is None else prelu_kwargs
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.use_transpose = use_transpose
        self.act_type = act_type
        self.antialiasing = use_antialiasing
        self.bias = None
        if self.antialiasing:
            if bias:
                self.bias = torch.nn.Parameter(torch.zeros(out_channels))
            self.low_pass_filter = BinomialAntiAlias(kernel_size=2 * kernel_size + 1)
            bias = False
        if act_type in ("snake", "snakebeta"):
            from ..bigvgan import AliasFreeSnake
            self.prelu = AliasFreeSnake(in_channels, alpha_logscale=True,
                                        beta=(act_type == "snakebeta"))
        elif act_type == "prelu":
            self.prelu = torch.nn.PReLU(device=device, dtype=dtype, **prelu_kwargs)
        elif act_type == "none" or act_type is None:
            self.prelu = None
        else:
            raise ValueError("'act_type' should be one of [prelu | snake]")
        conv_cls = torch.nn.ConvTranspose1d if use_transpose else torch.nn.Conv1d
        self.conv = conv_cls(in_channels, out_channels, kernel_size, stride=stride,
                             padding=padding, bias=bias, device=device, dtype=dtype)
        self.conv = cond_weight_norm(self.conv, use=use_weight_norm)

    def forward(self, x):
        return runtime.prelu_conv_forward(self, x)


class ConvBlock(torch.nn.Module):
    """UNIVERSE convolution block (paper appendix D; reference blocks.py:230-412)."""

    def __init__(self, n_channels, rate_change=None, rate_change_dir="none", act_type="prelu",
                 antialiasing=False, use_weight_norm=False, signal_cond_type=None):
        super().__init__()
        if rate_change_dir not in ["up", "down", "none"]:
            raise ValueError("The rate_change_dir value should be one of 'up' or 'down'")
        if rate_change_dir in ["up", "down"] and rate_change is None:
            raise ValueError("The rate_change should be specified when using for down/upsampling")
        if act_type != "prelu":
            raise NotImplementedError("only act_type='prelu' is on the shipped-config hot path")
        if signal_cond_type not in (None, "none"):
            raise NotImplementedError("signal_cond_type is dead code upstream (blocks.py:314-317)")
        self.rate = rate_change
        self.rate_change_dir = rate_change_dir
        self.n_channels = n_channels
        if rate_change_dir == "down":
            self.in_channels, self.out_channels = n_channels, 2 * n_channels
            self.rate_change_conv = PReLU_Conv(n_channels, 2 * n_channels, kernel_size=rate_change,
                                               stride=rate_change, use_weight_norm=use_weight_norm,
                                               use_antialiasing=antialiasing)
        elif rate_change_dir == "up":
            self.in_channels, self.out_channels = 2 * n_channels, n_channels
            self.rate_change_conv = PReLU_Conv(2 * n_channels, n_channels, kernel_size=rate_change,
                                               stride=rate_change, use_transpose=True,
                                               use_weight_norm=use_weight_norm,
                                               use_antialiasing=antialiasing)
        else:
            self.in_channels = self.out_channels = n_channels
            self.rate_change_conv = None
        self.conv1 = PReLU_Conv(n_channels, n_channels, kernel_size=5, padding="same",
                                act_type=act_type, use_weight_norm=use_weight_norm)
        self.conv2 = PReLU_Conv(n_channels, n_channels, kernel_size=3, padding="same",
                                act_type=act_type, use_weight_norm=use_weight_norm)
        self.conv3 = PReLU_Conv(n_channels, n_channels, kernel_size=3, padding="same",
                                act_type=act_type, use_weight_norm=use_weight_norm)
        self.signal_cond_proj = None

    def forward(self, h: torch.Tensor, noise_cond: Optional[torch.Tensor] = None,
                input_cond: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
                length: Optional[int] = None):
        """(B, C, T) fp32 in / out; returns (h_out, skip, cond_out) like the reference."""
        if res is not None and self.rate_change_dir == "down":
            raise ValueError("The residual input is not allowed for downsampling blocks")
        return runtime.conv_block_forward(self, h, noise_cond, input_cond, res, length)
