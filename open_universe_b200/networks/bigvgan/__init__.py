"""Inference-reachable part of the reference's ``networks/bigvgan`` package.

Upstream this package holds the BigVGAN discriminators / GAN losses (training only, out of
scope -- SURVEY.md section 2 #10) and the Snake activations.  Only ``AliasFreeSnake`` can be
reached from ``enhance()``, through ``UniverseGAN.signal_decoupling_layer`` when
``use_aux_signal`` / ``warm_start`` is requested (universe.py:317-331).  The modules keep the
reference's ``state_dict`` keys (``act.act.alpha``, ``act.upsample.kernel``,
``act.downsample.kernel``); the arithmetic is one CUDA kernel (``csrc/snake.cu``,
``ou_alias_free_snake``) which reads the resampling taps from those buffers.
"""
import torch
import torchaudio


class Snake(torch.nn.Module):
    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.alpha = torch.nn.Parameter(init * alpha, requires_grad=alpha_trainable)


class SnakeBeta(Snake):
    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__(in_features, alpha, alpha_trainable, alpha_logscale)
        init = torch.zeros(in_features) if alpha_logscale else torch.ones(in_features)
        self.beta = torch.nn.Parameter(init * alpha, requires_grad=alpha_trainable)


class Activation1d(torch.nn.Module):
    """x2 sinc up-sampling -> activation -> x2 down-sampling (alias_free_act.py:8-30)."""

    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2):
        super().__init__()
        self.up_ratio, self.down_ratio = up_ratio, down_ratio
        self.act = activation
        self.upsample = torchaudio.transforms.Resample(orig_freq=1, new_freq=up_ratio)
        self.downsample = torchaudio.transforms.Resample(orig_freq=up_ratio, new_freq=1)


class AliasFreeSnake(torch.nn.Module):
    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False,
                 beta=False):
        super().__init__()
        cls = SnakeBeta if beta else Snake
        self.act = Activation1d(cls(in_features, alpha=alpha, alpha_trainable=alpha_trainable,
                                    alpha_logscale=alpha_logscale))

    def forward(self, x):
        from ...engine import runtime
        return runtime.alias_free_snake(self, x)
