"""The UNIVERSE score network -- module surface of the reference's ``networks/universe/score.py``.

``ScoreNetwork.forward(x, sigma, cond)`` (score.py:277-297) evaluates the strided-conv U-Net
with FiLM sigma conditioning, a bidirectional GRU at the bottleneck (score.py:82-89,116) and
per-level signal conditioning.  Here the modules only hold parameters under the reference's
names; the forward is lowered once to a list of fused CUDA launches
(``engine.program.lower_score_network``) and replayed by ``engine.runtime``.
"""
import torch

from ...config import instantiate
from ...engine import runtime
from .blocks import ConvBlock, PReLU_Conv, cond_weight_norm
from .sigma_block import SigmaBlock, SimpleTimeEmbedding


class ScoreEncoder(torch.nn.Module):
    def __init__(self, ds_factors, input_channels, noise_cond_dim, with_gru_conv_sandwich=False,
                 with_extra_conv_block=False, act_type="prelu", use_weight_norm=False,
                 seq_model="gru", use_antialiasing=False):
        super().__init__()
        c = input_channels
        self.extra_conv_block = with_extra_conv_block
        self.ds_modules = torch.nn.ModuleList([
            ConvBlock(c * 2**i, r, "down", act_type=act_type, use_weight_norm=use_weight_norm,
                      antialiasing=use_antialiasing)
            for i, r in enumerate(ds_factors)])
        self.cond_proj = torch.nn.ModuleList([
            cond_weight_norm(torch.nn.Linear(noise_cond_dim, c * 2 ** (i + 1)), use=use_weight_norm)
            for i in range(len(ds_factors))])
        oc = input_channels * 2 ** len(ds_factors)
        if self.extra_conv_block:
            self.ds_modules.append(ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm))
            self.cond_proj.append(
                cond_weight_norm(torch.nn.Linear(noise_cond_dim, 2 * oc), use=use_weight_norm))
        self.seq_model = seq_model
        self.gru_conv_sandwich = False
        if seq_model == "gru":
            self.gru = torch.nn.GRU(oc, oc // 2, num_layers=1, bidirectional=True, batch_first=True)
            self.gru_conv_sandwich = with_gru_conv_sandwich
            if self.gru_conv_sandwich:
                self.conv_block1 = ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm)
                self.conv_block2 = ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm)
        elif seq_model != "none":
            raise ValueError("Values for 'seq_model' can be gru|attention|none")


class ScoreDecoder(torch.nn.Module):
    def __init__(self, up_factors, input_channels, noise_cond_dim, with_extra_conv_block=False,
                 act_type="prelu", use_weight_norm=False, use_antialiasing=False):
        super().__init__()
        self.extra_conv_block = with_extra_conv_block
        n_up = len(up_factors)
        n_channels = [input_channels * 2 ** (n_up - i - 1) for i in range(n_up)]
        self.up_modules = torch.nn.ModuleList()
        self.noise_cond_proj = torch.nn.ModuleList()
        self.signal_cond_proj = torch.nn.ModuleList()

        def add_level(block, c):
            self.up_modules.append(block)
            self.noise_cond_proj.append(
                cond_weight_norm(torch.nn.Linear(noise_cond_dim, 2 * c), use=use_weight_norm))
            self.signal_cond_proj.append(
                cond_weight_norm(torch.nn.Conv1d(c, c, kernel_size=1), use=use_weight_norm))

        if self.extra_conv_block:
            oc = input_channels * 2**n_up
            add_level(ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm), oc)
        for c, r in zip(n_channels, up_factors):
            add_level(ConvBlock(c, r, "up", act_type=act_type, use_weight_norm=use_weight_norm,
                                antialiasing=use_antialiasing), c)


class ScoreNetwork(torch.nn.Module):
    def __init__(self, fb_kernel_size=3, rate_factors=[2, 4, 4, 5], n_channels=32, n_rff=32,
                 noise_cond_dim=512, encoder_gru_conv_sandwich=False, extra_conv_block=False,
                 encoder_act_type="prelu", decoder_act_type="prelu", precoding=None,
                 input_channels=1, output_channels=1, use_weight_norm=False, seq_model="gru",
                 use_antialiasing=False, time_embedding=None):
        super().__init__()
        rate_factors = list(rate_factors)
        if input_channels != 1 or output_channels != 1:
            raise NotImplementedError("multi-channel signals are not used by any shipped config")
        if time_embedding == "simple":
            self.sigma_block = SimpleTimeEmbedding(n_dim=noise_cond_dim)
        else:
            self.sigma_block = SigmaBlock(n_rff, noise_cond_dim)
        self.input_channels, self.output_channels = input_channels, output_channels
        self.noise_cond_dim = noise_cond_dim
        self.rate_factors = rate_factors
        self.input_conv = torch.nn.Conv1d(input_channels, n_channels, kernel_size=fb_kernel_size,
                                          padding="same")
        self.encoder = ScoreEncoder(ds_factors=rate_factors, input_channels=n_channels,
                                    noise_cond_dim=noise_cond_dim,
                                    with_gru_conv_sandwich=encoder_gru_conv_sandwich,
                                    with_extra_conv_block=extra_conv_block,
                                    act_type=encoder_act_type, use_weight_norm=use_weight_norm,
                                    seq_model=seq_model, use_antialiasing=use_antialiasing)
        self.decoder = ScoreDecoder(up_factors=rate_factors[::-1], input_channels=n_channels,
                                    noise_cond_dim=noise_cond_dim,
                                    with_extra_conv_block=extra_conv_block,
                                    act_type=decoder_act_type, use_weight_norm=use_weight_norm,
                                    use_antialiasing=use_antialiasing)
        self.prelu = torch.nn.PReLU()
        self.output_conv = PReLU_Conv(n_channels, output_channels, kernel_size=fb_kernel_size,
                                      padding="same", use_weight_norm=use_weight_norm)
        self.precoding = instantiate(precoding, _recursive_=True) if precoding else None

    def forward(self, x, sigma, cond):
        """x (B,1,T) fp32, sigma (B,), cond: list of (B,C_l,T_l) fp32, coarsest first -> (B,1,T)."""
        return runtime.score_forward(self, x, sigma, cond)
