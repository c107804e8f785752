"""GPU sample-rate conversion for the CLI (reference bin/enhance.py:77-80,188-190 calls
``torchaudio.functional.resample``; torchaudio is a third-party dependency of the reference, its
default algorithm -- "sinc_interp_hann", lowpass_filter_width = 6, rolloff = 0.99 -- is restated here).

The polyphase filter bank is built on the host in float64; the convolution runs in ``ou_resample_poly``."""
import math

import torch

from ..engine import lib, runtime

_KERNELS = {}


def sinc_resample_kernel(orig, new, lowpass_filter_width=6, rolloff=0.99):
    """(kernel fp32 [new][2 * width + orig], width) for rates already divided by their gcd: for output phase
    p, tap j weighs input sample n * orig + j - width (torchaudio ``_get_sinc_resample_kernel``)."""
    key = (orig, new, lowpass_filter_width, rolloff)
    if key not in _KERNELS:
        base_freq = min(orig, new) * rolloff
        width = math.ceil(lowpass_filter_width * orig / base_freq)
        idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, :] / orig
        t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None] / new + idx
        t = (t * base_freq).clamp(-lowpass_filter_width, lowpass_filter_width)
        window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
        t = t * math.pi
        scale = base_freq / orig
        k = torch.where(t == 0, torch.ones_like(t), t.sin() / t) * window * scale
        _KERNELS[key] = (k.float().contiguous(), width)
    return _KERNELS[key]


@runtime.on_tensor_device
def resample(audio, fs, target_fs):
    """(..., T) fp32 CUDA tensor at ``fs`` -> (..., ceil(target_fs * T / fs)) at ``target_fs``."""
    fs, target_fs = int(fs), int(target_fs)
    if fs == target_fs:
        return audio
    runtime.require_cuda(audio)
    g = math.gcd(fs, target_fs)
    orig, new = fs // g, target_fs // g
    kern, width = sinc_resample_kernel(orig, new)
    shape = audio.shape
    x = audio.reshape(-1, shape[-1]).contiguous().float()
    t_in = x.shape[-1]
    t_out = math.ceil(new * t_in / orig)
    out = torch.empty(x.shape[0], t_out, dtype=torch.float32, device=x.device)
    kern = kern.to(x.device)
    lib.check(lib.load().ou_resample_poly(runtime._ptr(x), runtime._ptr(kern), runtime._ptr(out), x.shape[0],
                                          t_in, t_out, orig, new, width, kern.shape[1], runtime._stream()))
    return out.reshape(shape[:-1] + (t_out,))
