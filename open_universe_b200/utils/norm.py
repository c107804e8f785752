"""Batch amplitude normalisation (reference ``utils/norm.py:47-87``).

On CUDA tensors ``normalize_batch(norm=2, zero_mean=True)`` of a single signal runs in one fused
kernel (``ou_pad_normalize`` with zero padding); there is no CPU path.  The other norms
('max', '2-max') and target handling are training-side options that no shipped inference config
uses and are not built.
"""
import torch

from ..engine import lib, runtime


def normalize_batch(batch, norm=2, level_db=0.0, ref="noisy", eps=1e-5, zero_mean=True):
    assert ref in ["noisy", "both"]
    if norm not in (2, "2") or not zero_mean or eps != 1e-5:
        raise NotImplementedError("only norm=2 / zero_mean=True / eps=1e-5 is on the enhance() path")
    mix, *others = batch
    if any(t is not None for t in others):
        raise NotImplementedError("target normalisation is a training/debug option (not built)")
    runtime.require_cuda(mix)
    b, c, t = mix.shape
    if c != 1:
        raise NotImplementedError("multi-channel clips are not used by any shipped config")
    level = 10 ** (level_db / 20.0)
    src = mix.contiguous().float()
    out = torch.empty_like(src)
    stats = torch.empty(b, 2, dtype=torch.float32, device=mix.device)
    lib.check(lib.load().ou_pad_normalize(runtime._ptr(src), runtime._ptr(out), runtime._ptr(stats),
                                          b, t, t, 0, level, runtime._stream()))
    mean = stats[:, 0].reshape(b, 1, 1)
    inv_gain = stats[:, 1].reshape(b, 1, 1)
    return [out] + [None for _ in others], mean, inv_gain


def denormalize_batch(x, mean, std):
    return x * std + mean
