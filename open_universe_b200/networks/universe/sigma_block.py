"""Diffusion-time (sigma) embeddings -- parameter containers.

Reference: ``networks/universe/sigma_block.py``.  ``SigmaBlock`` :36-57 (UNIVERSE: random
Fourier features with a fixed ``freq`` buffer + three Linear/PReLU layers) and
``SimpleTimeEmbedding`` :60-78 (UNIVERSE++: sinusoid bank whose single frequency is a learned
sigmoid of log10 sigma).  The arithmetic runs in ``csrc`` (``ou_sigma_embed``).
"""
import torch

from ...engine import runtime


class Linear_PReLU(torch.nn.Module):
    def __init__(self, in_features, out_features, prelu_kwargs=None):
        super().__init__()
        self.prelu = torch.nn.PReLU(**({} if prelu_kwargs is None else prelu_kwargs))
        self.lin = torch.nn.Linear(in_features, out_features)


class SigmaBlock(torch.nn.Module):
    def __init__(self, n_rff=32, n_dim=256, scale=16):
        super().__init__()
        self.n_rff, self.n_dim = n_rff, n_dim
        self.register_buffer("freq", scale * torch.zeros(n_rff).normal_())
        self.layer1 = Linear_PReLU(2 * n_rff, 4 * n_rff)
        self.layer2 = Linear_PReLU(4 * n_rff, 8 * n_rff)
        self.layer3 = Linear_PReLU(8 * n_rff, n_dim)

    def forward(self, log10_sigma):
        return runtime.sigma_embed(self, log10_sigma)


class SimpleTimeEmbedding(torch.nn.Module):
    def __init__(self, n_dim=256):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.zeros((1, 1)), requires_grad=True)
        self.bias = torch.nn.Parameter(torch.zeros((1, 1)), requires_grad=True)
        self.n_dim = n_dim

    def forward(self, log10_sigma):
        return runtime.sigma_embed(self, log10_sigma)
