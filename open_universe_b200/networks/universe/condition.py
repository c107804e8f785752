"""The UNIVERSE conditioner network -- module surface of the reference's
``networks/universe/condition.py``.

``ConditionerNetwork.forward(x, x_wav=None, train=False)`` (condition.py:346-377): mel adapter
(STFT -> HTK mel -> global energy norm -> conv -> ConvBlock, :68-114), strided encoder with three
big ``st_convs`` into the latent (:33-65, :189-220), two ConvBlocks around a 2-layer BiGRU, and
a decoder whose per-block ``conv1`` outputs are the conditioning tensors (:264-270).  Runs once
per ``enhance()`` call.  Modules hold parameters; ``engine.program.lower_conditioner`` lowers it.
"""
import math

import torch
import torchaudio

from ...config import instantiate
from ...engine import runtime
from .blocks import BinomialAntiAlias, ConvBlock, PReLU_Conv, cond_weight_norm


def make_st_convs(ds_factors, input_channels, num_layers=None, use_weight_norm=False,
                  use_antialiasing=False):
    """Strided convs taking encoder level i straight to the latent rate: kernel = stride =
    prod(ds_factors[i:]) (condition.py:33-65)."""
    n = len(ds_factors)
    num_layers = n - 1 if num_layers is None else num_layers
    if use_antialiasing:
        raise NotImplementedError("anti-aliased st_convs are never built upstream (condition.py:333)")
    out_ch = input_channels * 2**n
    st_convs = torch.nn.ModuleList()
    for i in range(n):
        if i >= num_layers:
            st_convs.append(None)
            continue
        rate = math.prod(ds_factors[i:])
        st_convs.append(PReLU_Conv(input_channels * 2**i, out_ch, kernel_size=rate, stride=rate,
                                   use_weight_norm=use_weight_norm))
    return st_convs


class MelAdapter(torch.nn.Module):
    def __init__(self, n_mels, output_channels, ds_factor, oversample=2, use_weight_norm=False):
        super().__init__()
        self.ds_factor = ds_factor
        self.n_mels = n_mels
        self.n_fft = n_fft = oversample * ds_factor
        # sample_rate=24000 whatever the model's fs: it only shapes the mel filterbank, and the
        # reference hard-codes it (condition.py:76)
        self.mel_spec = torchaudio.transforms.MelSpectrogram(
            sample_rate=24000, n_mels=n_mels, n_fft=n_fft, hop_length=ds_factor, center=False)
        self.conv = cond_weight_norm(
            torch.nn.Conv1d(n_mels, output_channels, kernel_size=3, padding="same"),
            use=use_weight_norm)
        self.conv_block = ConvBlock(output_channels, use_weight_norm=use_weight_norm)
        pad_tot = n_fft - ds_factor
        self.pad_left, self.pad_right = pad_tot // 2, pad_tot - pad_tot // 2

    def compute_mel_spec(self, x):
        """(B,1,T) -> (B, n_mels, ceil(T/hop)) energy-normalised mel power spectrogram
        (condition.py:92-108); frame m covers samples [hop*m - pad_left, hop*m - pad_left + n_fft)."""
        return runtime.compute_mel_spec(self, x)


class ConditionerEncoder(torch.nn.Module):
    def __init__(self, ds_factors, input_channels, with_gru_residual=False,
                 with_extra_conv_block=False, act_type="prelu", use_weight_norm=False,
                 seq_model="gru", use_antialiasing=False):
        super().__init__()
        self.with_gru_residual = with_gru_residual
        self.extra_conv_block = with_extra_conv_block
        c = input_channels
        self.ds_modules = torch.nn.ModuleList([
            ConvBlock(c * 2**i, r, "down", act_type=act_type, use_weight_norm=use_weight_norm,
                      antialiasing=use_antialiasing)
            for i, r in enumerate(ds_factors)])
        self.st_convs = make_st_convs(ds_factors, input_channels, num_layers=len(ds_factors) - 1,
                                      use_weight_norm=use_weight_norm,
                                      use_antialiasing=use_antialiasing)
        oc = input_channels * 2 ** len(ds_factors)
        if self.extra_conv_block:
            self.ds_modules.append(ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm))
            self.st_convs.append(None)
        self.seq_model = seq_model
        if seq_model == "gru":
            self.gru = torch.nn.GRU(oc, oc // 2, num_layers=2, bidirectional=True, batch_first=True)
            self.conv_block1 = ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm)
            self.conv_block2 = ConvBlock(oc, act_type=act_type, use_weight_norm=use_weight_norm)
        else:
            raise ValueError("Values for 'seq_model' can be gru|attention")


class ConditionerDecoder(torch.nn.Module):
    def __init__(self, up_factors, input_channels, with_extra_conv_block=False, act_type="prelu",
                 use_weight_norm=False, use_antialiasing=False):
        super().__init__()
        self.extra_conv_block = with_extra_conv_block
        n_up = len(up_factors)
        n_channels = [input_channels * 2 ** (n_up - i - 1) for i in range(n_up)]
        self.input_conv_block = ConvBlock(n_channels[0] * 2, act_type=act_type,
                                          use_weight_norm=use_weight_norm)
        blocks = []
        if self.extra_conv_block:
            blocks.append(ConvBlock(2 * n_channels[0], act_type=act_type,
                                    use_weight_norm=use_weight_norm))
        blocks += [ConvBlock(c, r, "up", act_type=act_type, use_weight_norm=use_weight_norm,
                             antialiasing=use_antialiasing)
                   for c, r in zip(n_channels, up_factors)]
        self.up_modules = torch.nn.ModuleList(blocks)


class ConditionerNetwork(torch.nn.Module):
    def __init__(self, fb_kernel_size=3, rate_factors=[2, 4, 4, 5], n_channels=32, n_mels=80,
                 n_mel_oversample=4, encoder_gru_residual=False, extra_conv_block=False,
                 encoder_act_type="prelu", decoder_act_type="prelu", precoding=None,
                 input_channels=1, output_channels=None, use_weight_norm=False, seq_model="gru",
                 use_antialiasing=False):
        super().__init__()
        rate_factors = list(rate_factors)
        if input_channels != 1:
            raise NotImplementedError("multi-channel signals are not used by any shipped config")
        self.n_mels = n_mels
        self.rate_factors = rate_factors
        self.input_conv = cond_weight_norm(
            torch.nn.Conv1d(input_channels, n_channels, kernel_size=fb_kernel_size, padding="same"),
            use=use_weight_norm)
        if output_channels is not None:
            self.output_conv = cond_weight_norm(
                torch.nn.Conv1d(n_channels, output_channels, kernel_size=fb_kernel_size,
                                padding="same"), use=use_weight_norm)
        else:
            self.output_conv = None
        total_ds = math.prod(rate_factors)
        total_channels = 2 ** len(rate_factors) * n_channels
        self.input_mel = MelAdapter(n_mels, total_channels, total_ds * input_channels,
                                    n_mel_oversample, use_weight_norm=use_weight_norm)
        # the encoder never uses anti-aliasing upstream (condition.py:333)
        self.encoder = ConditionerEncoder(rate_factors, n_channels,
                                          with_gru_residual=encoder_gru_residual,
                                          with_extra_conv_block=extra_conv_block,
                                          act_type=encoder_act_type,
                                          use_weight_norm=use_weight_norm, seq_model=seq_model,
                                          use_antialiasing=False)
        self.decoder = ConditionerDecoder(rate_factors[::-1], n_channels,
                                          with_extra_conv_block=extra_conv_block,
                                          act_type=decoder_act_type,
                                          use_weight_norm=use_weight_norm,
                                          use_antialiasing=use_antialiasing)
        self.precoding = instantiate(precoding, _recursive_=True) if precoding else None

    def forward(self, x, x_wav=None, train=False):
        """(B,1,T) -> conditions (list, coarsest first) [, y_hat, h when train=True]."""
        conditions, y_hat, h = runtime.conditioner_forward(self, x, x_wav)
        if train:
            return conditions, y_hat, h
        return conditions
