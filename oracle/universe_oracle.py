"""CPU oracle: a plain-PyTorch fp32 restatement of open-universe's ``enhance()`` hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file, and only as
the checker / the timed CPU baseline.  The product (``open_universe_b200``) never imports it
and has no CPU fallback.

Parity status: PINNED.  The reference has no tests or golden vectors of its own (SURVEY.md
section 4), so this restatement is pinned against outputs of the unmodified reference run in
the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); see
``tests/test_oracle_golden.py``.

Everything is a pure function of ``(cfg, sd)``: ``cfg`` the resolved ``model:`` mapping of a
reference YAML (``config/model/*.yaml``) and ``sd`` a ``state_dict`` in the reference's own key
layout (weight-norm stored as ``weight_g`` / ``weight_v``).  Citations are ``file:line`` under
``/root/reference/open_universe/``.

Third-party arithmetic used by the reference on this path, restated with the same library
calls because they are importable everywhere the oracle runs: ``torch.nn.functional.conv1d``,
``conv_transpose1d``, ``prelu``, ``torch.nn.GRU`` (gate semantics additionally restated
explicitly in ``gru_explicit``) and ``torch.fft.rfft`` (torchaudio ``Spectrogram`` ==
``torch.stft`` with a periodic Hann window, ``center=False``, ``power=2``).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT_HALF = 1.0 / math.sqrt(2.0)


# ----------------------------------------------------------------------------- helpers
def binomial_taps(kernel_size):
    """networks/universe/blocks.py:62-68 -- Pascal row scaled to unit RMS (not unit sum)."""
    row = np.array([math.comb(kernel_size - 1, i) for i in range(kernel_size)], dtype=np.float64)
    full = np.zeros((kernel_size, kernel_size))
    for n in range(kernel_size):  # lower-triangular Pascal matrix, as scipy.linalg.pascal
        for k in range(n + 1):
            full[n, k] = math.comb(n, k)
    norm = np.sqrt(np.mean(full**2))
    w = torch.tensor((row / norm).astype("float32"), dtype=torch.float32)
    return w / w.square().mean().sqrt()


def eff_weight(sd, prefix):
    """Old-style ``torch.nn.utils.weight_norm(dim=0)``: w = g * v / ||v||, norm over all dims
    but 0 (blocks.py:36-41).  dim 0 is C_out for Conv1d/Linear and C_in for ConvTranspose1d."""
    if prefix + ".weight_g" in sd:
        g, v = sd[prefix + ".weight_g"], sd[prefix + ".weight_v"]
        dims = tuple(range(1, v.ndim))
        return v * (g / v.norm(2, dim=dims, keepdim=True))
    return sd[prefix + ".weight"]


def lowpass(x, taps):
    """blocks.py:119-130 -- depthwise FIR, zero 'same' padding."""
    c = x.shape[1]
    w = taps.to(x)[None, None, :].expand(c, 1, -1)
    return F.conv1d(x, w, padding="same", groups=c)


def film(x, y):
    """blocks.py:53-59."""
    if y.shape[1] != 2 * x.shape[1]:
        raise ValueError("g should have 2 times more channels than y")
    c = x.shape[1]
    return y[:, :c, None] * x + y[:, c:, None]


def prelu_conv(sd, p, x, stride=1, transpose=False, same=False, antialias=False):
    """blocks.py:205-227 (PReLU_Conv.forward) with act_type='prelu'."""
    r = x.shape[-1] % stride
    if not transpose and r != 0:
        x = F.pad(x, (0, stride - r))
    x = F.prelu(x, sd[p + ".prelu.weight"])
    taps = sd.get(p + ".low_pass_filter.weights") if antialias else None
    if antialias and not transpose:
        x = lowpass(x, taps)
    w = eff_weight(sd, p + ".conv")
    b = sd.get(p + ".conv.bias")
    if transpose:
        x = F.conv_transpose1d(x, w, b, stride=stride)
    else:
        x = F.conv1d(x, w, b, stride=stride, padding="same" if same else 0)
    if antialias and transpose:
        x = lowpass(x, taps)
    if antialias and (p + ".bias") in sd:
        x = x + sd[p + ".bias"][None, :, None]
    return x


def conv_block(sd, p, h, direction="none", rate=None, antialias=False, noise_cond=None,
               input_cond=None, res=None, length=None):
    """blocks.py:327-412 (ConvBlock.forward).  Returns (h_out, skip, cond_out)."""
    if direction == "up":
        if length is not None and rate * h.shape[-1] < length:
            h = F.pad(h, (0, 1))
        h = prelu_conv(sd, p + ".rate_change_conv", h, stride=rate, transpose=True,
                       antialias=antialias)
        if length is not None:
            h = F.pad(h, (0, length - h.shape[-1]))
    if res is not None:
        h = (h + res) * SQRT_HALF
    cond_out = prelu_conv(sd, p + ".conv1", h, same=True)
    c = cond_out
    if input_cond is not None:
        c = (cond_out + input_cond) * SQRT_HALF
    if noise_cond is not None:
        c = film(c, noise_cond)
    c = prelu_conv(sd, p + ".conv2", c, same=True)
    c = prelu_conv(sd, p + ".conv3", c, same=True)
    v = (h + c) * SQRT_HALF
    if direction == "down":
        r = h.shape[-1] % rate
        v_pad = F.pad(v, (0, rate - r)) if r != 0 else v
        h = prelu_conv(sd, p + ".rate_change_conv", v_pad, stride=rate, antialias=antialias)
        return h, v, cond_out
    return v, v, cond_out


def linear(sd, p, x):
    return F.linear(x, eff_weight(sd, p), sd.get(p + ".bias"))


def gru_explicit(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """Explicit restatement of one direction of ``torch.nn.GRU`` (h0 = 0), gates ordered
    (r, z, n):  r = s(Wir x + bir + Whr h + bhr), z likewise, n = tanh(Win x + bin + r*(Whn h + bhn)),
    h' = (1 - z) * n + z * h.  x: (B, T, I) -> (B, T, H).  This is the arithmetic the CUDA
    recurrence kernel implements (score.py:82-89,116)."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    xp = x @ w_ih.t() + b_ih
    h = x.new_zeros(B, H)
    out = x.new_zeros(B, T, H)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        hp = h @ w_hh.t() + b_hh
        r = torch.sigmoid(xp[:, t, :H] + hp[:, :H])
        z = torch.sigmoid(xp[:, t, H:2 * H] + hp[:, H:2 * H])
        n = torch.tanh(xp[:, t, 2 * H:] + r * hp[:, 2 * H:])
        h = (1.0 - z) * n + z * h
        out[:, t] = h
    return out


_gru_cache = {}


def bigru(sd, p, x, num_layers=1):
    """(B, C, T) -> (B, C, T) through torch.nn.GRU(C, C//2, bidirectional, batch_first)
    exactly as score.py:82-89,116-117 / condition.py:173-179,213."""
    c = x.shape[1]
    key = (p, num_layers, c, id(sd))
    gru = _gru_cache.get(key)
    if gru is None:
        gru = torch.nn.GRU(c, c // 2, num_layers=num_layers, bidirectional=True, batch_first=True)
        gru.load_state_dict({k[len(p) + 1:]: v for k, v in sd.items() if k.startswith(p + ".")})
        gru.eval()
        _gru_cache.clear()
        _gru_cache[key] = gru
    gru = gru.to(x.device)
    y, _ = gru(x.transpose(-2, -1))
    return y.transpose(-2, -1)


# ----------------------------------------------------------------------------- score network
def sigma_embedding(cfg_score, sd, p, log10_sigma):
    """sigma_block.py:60-78 (SimpleTimeEmbedding, UNIVERSE++) / :36-57 (SigmaBlock, UNIVERSE)."""
    n_dim = cfg_score.get("noise_cond_dim", 512)
    if cfg_score.get("time_embedding") == "simple":
        k = torch.arange(n_dim // 2, device=log10_sigma.device)
        f = 0.5 * torch.sigmoid(sd[p + ".weight"] * log10_sigma[:, None] + sd[p + ".bias"])
        ph = 2.0 * math.pi * f * k
        return torch.cat([torch.sin(ph), torch.cos(ph)], dim=-1)
    ph = 2.0 * math.pi * sd[p + ".freq"][None, :] * log10_sigma[:, None]
    g = torch.cat([torch.sin(ph), torch.cos(ph)], dim=-1)
    for layer in ("layer1", "layer2", "layer3"):
        g = F.prelu(linear(sd, f"{p}.{layer}.lin", g), sd[f"{p}.{layer}.prelu.weight"])
    return g


def score_network(cfg_score, sd, p, x, sigma, cond):
    """score.py:277-297 (ScoreNetwork.forward) with ScoreEncoder :104-127, ScoreDecoder :196-210."""
    rates = list(cfg_score.get("rate_factors", [2, 4, 4, 5]))
    extra = cfg_score.get("extra_conv_block", False)
    aa = cfg_score.get("use_antialiasing", False)
    seq_model = cfg_score.get("seq_model", "gru")
    n_samples = x.shape[-1]
    g = sigma_embedding(cfg_score, sd, p + ".sigma_block", torch.log10(sigma))
    x = F.conv1d(x, sd[p + ".input_conv.weight"], sd[p + ".input_conv.bias"], padding="same")
    residuals, lengths = [], []
    n_blocks = len(rates) + (1 if extra else 0)
    for i in range(n_blocks):
        nc = linear(sd, f"{p}.encoder.cond_proj.{i}", g)
        lengths.append(x.shape[-1])
        if i < len(rates):
            x, res, _ = conv_block(sd, f"{p}.encoder.ds_modules.{i}", x, "down", rates[i], aa,
                                   noise_cond=nc)
        else:
            x, res, _ = conv_block(sd, f"{p}.encoder.ds_modules.{i}", x, noise_cond=nc)
        residuals.append(res)
    if seq_model == "gru":
        if cfg_score.get("encoder_gru_conv_sandwich", False):
            x, *_ = conv_block(sd, p + ".encoder.conv_block1", x)
        x = bigru(sd, p + ".encoder.gru", x)
        if cfg_score.get("encoder_gru_conv_sandwich", False):
            x, *_ = conv_block(sd, p + ".encoder.conv_block2", x)
    residuals, lengths = residuals[::-1], lengths[::-1]
    up_rates = ([None] if extra else []) + rates[::-1]
    for lvl, r in enumerate(up_rates):
        nc = linear(sd, f"{p}.decoder.noise_cond_proj.{lvl}", g)
        sc = F.conv1d(cond[lvl], eff_weight(sd, f"{p}.decoder.signal_cond_proj.{lvl}"),
                      sd[f"{p}.decoder.signal_cond_proj.{lvl}.bias"])
        x, *_ = conv_block(sd, f"{p}.decoder.up_modules.{lvl}", x,
                           "none" if r is None else "up", r, aa, noise_cond=nc, input_cond=sc,
                           res=residuals[lvl], length=lengths[lvl])
    x = F.prelu(x, sd[p + ".prelu.weight"])
    x = prelu_conv(sd, p + ".output_conv", x, same=True)
    return F.pad(x, (0, n_samples - x.shape[-1]))


# ----------------------------------------------------------------------------- conditioner
def hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs, n_mels, sample_rate=24000, f_min=0.0, f_max=None):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') as instantiated at
    condition.py:75-81 -- NOTE sample_rate is 24000 whatever the model's fs."""
    f_max = sample_rate / 2 if f_max is None else f_max
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)  # (n_freqs, n_mels)


def mel_spec(cfg_cond, x, sd=None, p=None):
    """condition.py:92-108 (MelAdapter.compute_mel_spec).  x: (B, 1, T) -> (B, n_mels, frames).
    frame m covers samples [hop*m - pad_left, hop*m - pad_left + n_fft)."""
    rates = list(cfg_cond.get("rate_factors", [2, 4, 4, 5]))
    hop = math.prod(rates) * cfg_cond.get("input_channels", 1)
    n_fft = cfg_cond.get("n_mel_oversample", 4) * hop
    n_mels = cfg_cond.get("n_mels", 80)
    pad_tot = n_fft - hop
    pad_l, pad_r = pad_tot // 2, pad_tot - pad_tot // 2
    r = x.shape[-1] % hop
    pad = hop - r if r != 0 else 0
    x = F.pad(x, (pad_l, pad + pad_r))
    if sd is not None and (p + ".mel_spec.spectrogram.window") in sd:
        window = sd[p + ".mel_spec.spectrogram.window"]
        fb = sd[p + ".mel_spec.mel_scale.fb"]
    else:
        window = torch.hann_window(n_fft, periodic=True)
        fb = mel_filterbank(n_fft // 2 + 1, n_mels)
    frames = x.squeeze(1).unfold(-1, n_fft, hop)  # (B, frames, n_fft)
    spec = torch.fft.rfft(frames * window.to(x), dim=-1)
    power = spec.real.square() + spec.imag.square()  # (B, frames, n_freqs)
    mel = (power @ fb.to(x)).transpose(-2, -1)  # (B, n_mels, frames)
    norm = mel.square().sum(dim=-2, keepdim=True).mean(dim=-1, keepdim=True).sqrt()
    return mel / norm.clamp(min=1e-5)


def st_conv_rates(rates):
    """condition.py:42-46."""
    out = [rates[-1]]
    for r in rates[-2::-1]:
        out.append(out[-1] * r)
    return out[::-1]


def conditioner_network(cfg_cond, sd, p, x, x_wav=None):
    """condition.py:346-377 with train=True.  Returns (conditions[5], y_hat, h)."""
    rates = list(cfg_cond.get("rate_factors", [2, 4, 4, 5]))
    extra = cfg_cond.get("extra_conv_block", False)
    aa_dec = cfg_cond.get("use_antialiasing", False)  # encoder is hard-coded off (condition.py:333)
    n_samples = x.shape[-1]
    x_wav = x if x_wav is None else x_wav
    # MelAdapter.forward condition.py:110-114
    m = mel_spec(cfg_cond, x_wav, sd, p + ".input_mel")
    m = F.conv1d(m, eff_weight(sd, p + ".input_mel.conv"), sd[p + ".input_mel.conv.bias"],
                 padding="same")
    x_mel, *_ = conv_block(sd, p + ".input_mel.conv_block", m)
    x = F.conv1d(x, eff_weight(sd, p + ".input_conv"), sd[p + ".input_conv.bias"], padding="same")
    # ConditionerEncoder.forward condition.py:189-220
    st_rates = st_conv_rates(rates)
    outputs, lengths = [], []
    n_blocks = len(rates) + (1 if extra else 0)
    for i in range(n_blocks):
        lengths.append(x.shape[-1])
        if i < len(rates):
            x, res, _ = conv_block(sd, f"{p}.encoder.ds_modules.{i}", x, "down", rates[i], False)
        else:
            x, res, _ = conv_block(sd, f"{p}.encoder.ds_modules.{i}", x)
        if i < len(rates) - 1:
            outputs.append(prelu_conv(sd, f"{p}.encoder.st_convs.{i}", res, stride=st_rates[i]))
    outputs.append(x)
    out = x_mel
    for o in outputs:
        out = out + o
    out = out * (1.0 / math.sqrt(len(outputs) + 1))
    out, *_ = conv_block(sd, p + ".encoder.conv_block1", out)
    res = out
    out = bigru(sd, p + ".encoder.gru", out, num_layers=2)
    if cfg_cond.get("encoder_gru_residual", False):
        out = (out + res) / math.sqrt(2)
    h, *_ = conv_block(sd, p + ".encoder.conv_block2", out)
    lengths = lengths[::-1]
    # ConditionerDecoder.forward condition.py:264-270
    y, *_ = conv_block(sd, p + ".decoder.input_conv_block", h)
    conditions = []
    up_rates = ([None] if extra else []) + rates[::-1]
    for lvl, (r, length) in enumerate(zip(up_rates, lengths)):
        y, _, c = conv_block(sd, f"{p}.decoder.up_modules.{lvl}", y,
                             "none" if r is None else "up", r, aa_dec, length=length)
        conditions.append(c)
    if (p + ".output_conv.bias") in sd:
        y = F.conv1d(y, eff_weight(sd, p + ".output_conv"), sd[p + ".output_conv.bias"],
                     padding="same")
    y = F.pad(y, (0, n_samples - y.shape[-1]))
    return conditions, y, h


# ----------------------------------------------------------------------------- aux signal path
def sinc_resample(x, kernel, orig, new):
    """torchaudio.functional._apply_sinc_resample_kernel (torchaudio 2.x, third-party: not under
    /root/reference; call sites networks/bigvgan/alias_free_act.py:21-22,26-28) with the kernel
    taken from the module's ``kernel`` buffer, shape (new, 1, 2*width + orig):
    pad (width, width + orig) -> conv1d stride ``orig`` -> interleave the ``new`` phases ->
    crop to ceil(new * L / orig)."""
    b, c, length = x.shape
    width = (kernel.shape[-1] - orig) // 2
    y = F.pad(x.reshape(b * c, 1, length), (width, width + orig))
    y = F.conv1d(y, kernel.to(x), stride=orig)                # (b*c, new, frames)
    y = y.transpose(1, 2).reshape(b * c, -1)
    target = -(-new * length // orig)
    return y[:, :target].reshape(b, c, target)


def alias_free_snake(sd, p, x, beta=False):
    """networks/bigvgan/snake.py:127-157 + alias_free_act.py:25-30: x2 sinc up-sampling ->
    Snake with log-scale alpha (snake.py:52-62; SnakeBeta :116-124) -> /2 down-sampling."""
    x = sinc_resample(x, sd[p + ".act.upsample.kernel"], 1, 2)
    alpha = torch.exp(sd[p + ".act.act.alpha"])[None, :, None]
    mag = torch.exp(sd[p + ".act.act.beta"])[None, :, None] if beta else alpha
    x = x + (1.0 / (mag + 0.000000001)) * torch.sin(x * alpha) ** 2
    return sinc_resample(x, sd[p + ".act.downsample.kernel"], 2, 1)


def aux_to_wav(cfg, sd, y_aux):
    """universe_gan.py:117-126,145-149: PReLU_Conv(n_channels -> 1, k=3, 'same') whose activation
    is ``losses.signal_decoupling_act`` (snake in every shipped UNIVERSE++ config); identity for
    the original UNIVERSE class (universe.py has no decoupling layer)."""
    p = "signal_decoupling_layer"
    if (p + ".conv.weight") not in sd and (p + ".conv.weight_v") not in sd:
        return y_aux
    act = (cfg.get("losses") or {}).get("signal_decoupling_act", None)
    if act in ("snake", "snakebeta"):
        y_aux = alias_free_snake(sd, p + ".prelu", y_aux, beta=act == "snakebeta")
    elif act == "prelu":
        y_aux = F.prelu(y_aux, sd[p + ".prelu.weight"])
    elif act != "none":
        raise ValueError("'act_type' should be one of [prelu | snake]")
    return F.conv1d(y_aux, eff_weight(sd, p + ".conv"), sd.get(p + ".conv.bias"), padding="same")


def signal_median(signal):
    """utils/stats.py:22-66: pick, per batch row, the ensemble member that is the sample-wise
    median most often.  signal: (ensemble, batch, ...) -> (batch, ...)."""
    shape = signal.shape
    flat = signal.flatten(start_dim=2)
    n = flat.shape[0]
    _, order = flat.sort(dim=0)
    # rank position closest to n/2 holds the member that is "the median" of that sample; torch's
    # min() returns the first minimum, as upstream
    _, pos = (order - n / 2).abs().min(dim=0)                  # (batch, samples)
    counts = torch.stack([(pos == i).sum(dim=1) for i in range(n)], dim=1)
    select = counts.argmax(dim=1)
    out = torch.stack([flat[select[i], i] for i in range(flat.shape[1])], dim=0)
    return out.reshape(shape[1:])


# ----------------------------------------------------------------------------- enhance()
class UniverseOracle:
    """Functional stand-in for ``Universe`` / ``UniverseGAN`` inference (universe.py:231-375)."""

    def __init__(self, cfg, sd):
        self.cfg = cfg
        self.sd = {k: v.detach().float() for k, v in sd.items()}
        self.edm = cfg.get("edm")
        self.score_prefix = "_edm_model" if self.edm is not None else "score_model"
        self.rates = list(cfg["score_model"].get("rate_factors", [2, 4, 4, 5]))
        self.tot_ds = math.prod(self.rates)
        self.fs = cfg["fs"]

    def to(self, device):
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    # universe.py:219-226
    def pad(self, x, pad=None):
        if pad is None:
            pad = self.tot_ds - x.shape[-1] % self.tot_ds
        return F.pad(x, (pad // 2, pad - pad // 2)), pad

    def unpad(self, x, pad):
        return x[..., pad // 2: -(pad - pad // 2)]

    # utils/norm.py:47-87 with norm=2, ref='both', target=None
    def normalize(self, mix):
        kw = self.cfg.get("normalization_kwargs", {})
        level = 10.0 ** (kw.get("level_db", 0.0) / 20.0)
        mix = mix - mix.mean(dim=(1, 2), keepdim=True)
        gain = level / mix.std(dim=(1, 2), keepdim=True).clamp(min=1e-5)
        return mix * gain

    def condition(self, mix, x_wav=None):
        return conditioner_network(self.cfg["condition_model"], self.sd, "condition_model", mix,
                                   x_wav)

    def net(self, x, sigma, cond):
        return score_network(self.cfg["score_model"], self.sd, self.score_prefix, x, sigma, cond)

    # universe.py:175-209
    def score(self, x, sigma, cond):
        if self.edm is None:
            return self.net(x, sigma, cond)
        kw = self.cfg.get("normalization_kwargs", {})
        level_db = self.edm.get("data_level_db", kw.get("level_db", 0.0))
        sd_ = 10.0 ** (level_db / 20.0)
        s = sigma[:, None, None]
        s_norm = (s**2 + sd_**2) ** 0.5
        w_skip = sd_**2 / (s**2 + sd_**2)
        w_in = 1.0 / s_norm
        w_out = s * sd_ / s_norm
        net_out = self.net(w_in * x, self.edm["noise"] * sigma, cond)
        est = w_skip * x + w_out * net_out
        return (est - x) / s**2

    def sigmas(self, n_steps, like):
        """universe.py:308-311, 380-386: geometric schedule on linspace(0,1,N) flipped."""
        d = self.cfg["diffusion"]
        if d.get("schedule", "geometric") != "geometric":
            raise NotImplementedError()
        time = torch.linspace(0, 1, n_steps).type_as(like).flip(dims=[0])
        return float(d["sigma_min"]) * (float(d["sigma_max"]) / float(d["sigma_min"])) ** time

    def aux_to_wav(self, y_aux):
        return aux_to_wav(self.cfg, self.sd, y_aux)

    def normalize_target(self, mix_padded, target_padded):
        """utils/norm.py:72-86: ref='both' normalises the target on its own statistics, ref='noisy'
        with the mixture's mean and gain."""
        kw = self.cfg.get("normalization_kwargs", {})
        level = 10.0 ** (kw.get("level_db", 0.0) / 20.0)
        if kw.get("ref", "noisy") == "both":
            t = target_padded - target_padded.mean(dim=(1, 2), keepdim=True)
            return t * (level / t.std(dim=(1, 2), keepdim=True).clamp(min=1e-5))
        mean = mix_padded.mean(dim=(1, 2), keepdim=True)
        gain = level / (mix_padded - mean).std(dim=(1, 2), keepdim=True).clamp(min=1e-5)
        return (target_padded - mean) * gain

    def partial_diffusion(self, mix, n_steps, epsilon, t_final, noise):
        """UniverseLoRA.partial_diffusion (networks/universe/lora.py:231-296): ``n_steps`` sampler steps from
        t = 1 down to a per-clip final time ``t_final`` (B,), no padding, no post-processing.
        mix (B, 1, T); ``noise``: the n_steps unit-variance tensors in draw order."""
        d = self.cfg["diffusion"]
        ratio = float(d["sigma_max"]) / float(d["sigma_min"])
        delta_t = (1.0 - t_final) / (n_steps - 1)
        mix = self.normalize(mix)
        gamma = ratio ** -delta_t
        eta = 1 - gamma**epsilon
        beta = torch.sqrt(1 - gamma ** (2 * (epsilon - 1.0)))
        time = mix.new_ones(mix.shape[0])
        std = lambda t: float(d["sigma_min"]) * ratio ** t   # noqa: E731  (universe.py:380-386)
        sigma = std(time)
        cond, _, _ = self.condition(mix, x_wav=mix)
        it = iter(noise)
        x = next(it).to(mix) * sigma[:, None, None]
        for _ in range(n_steps - 1):
            score = self.score(x, sigma, cond)
            time = time - delta_t
            sigma_next = std(time)
            z = next(it).to(mix) * sigma_next[:, None, None]
            x = x + sigma[:, None, None] ** 2 * eta[:, None, None] * score + beta[:, None, None] * z
            sigma = sigma_next
        score = self.score(x, sigma, cond)
        return x + sigma[:, None, None] ** 2 * score

    def enhance(self, mix, n_steps=None, epsilon=None, noise=None, rng=None, keep_rms=False,
                use_aux_signal=False, ensemble=None, ensemble_stat="median", warm_start=None,
                target=None, fake_score_snr=None):
        """universe.py:231-375.  ``noise``: optional list of unit-variance (B,1,T_pad) tensors used
        in draw order instead of the module-level ``randn`` (universe.py:39-41); the oracle-score
        debugging noise of ``score_wrapper`` (:293-300) always comes from torch.randn, as upstream."""
        d = self.cfg["diffusion"]
        epsilon = d["epsilon"] if epsilon is None else epsilon
        n_steps = d["n_steps"] if n_steps is None else n_steps
        x_ndim = mix.ndim
        if x_ndim == 1:
            mix = mix[None, None, :]
        elif x_ndim == 2:
            mix = mix[:, None, :]
        elif x_ndim > 3:
            raise ValueError("The input should have at most 3 dimensions")
        mix_rms = mix.square().mean(dim=(-2, -1), keepdim=True).sqrt()
        if ensemble is not None:
            mix_shape = mix.shape
            mix = torch.stack([mix] * ensemble, dim=0).view((-1,) + mix_shape[1:])
        mix_len = mix.shape[-1]
        mix, pad = self.pad(mix)
        if target is not None:
            target, _ = self.pad(target, pad=pad)
            target = self.normalize_target(mix, target)
        mix = self.normalize(mix)
        score_snr = 5.0 if fake_score_snr is None else fake_score_snr
        it = iter(noise) if noise is not None else None

        def score_wrapper(x, s, cond):
            # universe.py:284-300
            if target is None:
                return self.score(x, s, cond)
            true_score = -(x - target) / s[:, None, None] ** 2
            noise_rms = (true_score**2).mean().sqrt() * 10 ** (-score_snr / 20.0)
            nz = torch.randn(true_score.shape, dtype=true_score.dtype, device=true_score.device,
                             generator=rng)
            return true_score + nz * noise_rms

        def randn(sig):
            if it is not None:
                n = next(it).to(mix)
            else:
                n = torch.randn(mix.shape, dtype=mix.dtype, device=mix.device, generator=rng)
            return n * sig[:, None, None]

        delta_t = 1.0 / (n_steps - 1)
        gamma = (float(d["sigma_max"]) / float(d["sigma_min"])) ** -delta_t
        eta = 1 - gamma**epsilon
        beta = math.sqrt(1 - gamma ** (2 * (epsilon - 1.0)))
        sigma = self.sigmas(n_steps, mix)
        sigma = sigma[None, :].expand(mix.shape[0], -1)
        cond, aux_signal, _ = self.condition(mix, x_wav=mix)
        if use_aux_signal:
            x = self.aux_to_wav(aux_signal)
        else:
            if warm_start is None:
                x = randn(sigma[:, 0])
                n_start = 0
            else:
                x = self.aux_to_wav(aux_signal) + randn(sigma[:, warm_start])
                n_start = warm_start
            for n in range(n_start, n_steps - 1):
                s_now, s_next = sigma[:, n], sigma[:, n + 1]
                score = score_wrapper(x, s_now, cond)
                z = randn(s_next)
                x = x + s_now[:, None, None] ** 2 * eta * score + beta * z
            score = score_wrapper(x, sigma[:, -1], cond)
            x = x + sigma[:, -1, None, None] ** 2 * score
        x = self.unpad(x, pad)
        x = F.pad(x, (0, mix_len - x.shape[-1]))
        if keep_rms:
            x_rms = x.square().mean(dim=(-2, -1), keepdim=True).sqrt().clamp(min=1e-5)
            x = x * (mix_rms / x_rms)
        scale = abs(x).max(dim=-1, keepdim=True).values
        x = torch.where(scale > 1.0, x / scale, x)
        if ensemble is not None:
            x = x.view((-1,) + mix_shape)
            if ensemble_stat == "mean":
                x = x.mean(dim=0)
            elif ensemble_stat == "median":
                x = x.median(dim=0).values
            elif ensemble_stat == "signal_median":
                x = signal_median(x)
            else:
                raise NotImplementedError()
        if x_ndim == 1:
            x = x[0, 0]
        elif x_ndim == 2:
            x = x[:, 0, :]
        return x
