"""Deterministic synthetic weights / inputs shared by the golden generator and the tests.

TEST INFRASTRUCTURE.  No checkpoint exists offline, so oracle, golden vectors and the CUDA
path all run on weights produced by this file from a (key -> shape) manifest of the
reference ``state_dict`` (SURVEY.md section 8c "Weights").  Only numpy's PCG64 streams are
used, so every box regenerates bit-identical tensors.  Buffers that the reference builds
deterministically in its constructors (binomial taps, Hann window, mel filterbank) are NOT
generated here: the product has to rebuild them itself and the tests compare them with the
values stored in the golden files.
"""
import zlib

import numpy as np
import torch

CONSTRUCTOR_BUFFERS = ("low_pass_filter.weights", "spectrogram.window", "mel_scale.fb")


def _rng(seed, key):
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def is_constructor_buffer(key):
    return key.endswith(CONSTRUCTOR_BUFFERS)


def det_tensor(key, shape, seed=0, v_for_g=None):
    shape = tuple(shape)
    r = _rng(seed, key)
    if key.endswith("weight_g"):
        # weight-norm gain: ||v|| over dims (1,2) perturbed by +-10 % (blocks.py:36-41)
        nrm = np.sqrt((v_for_g.astype(np.float64) ** 2).sum(axis=tuple(range(1, v_for_g.ndim)),
                                                             keepdims=True))
        out = nrm * (1.0 + 0.1 * r.uniform(-1, 1, size=nrm.shape))
    elif key.endswith("prelu.weight") or (key.endswith(".weight") and shape == (1,)):
        out = 0.25 + 0.1 * r.uniform(-1, 1, size=shape)
    elif key.endswith("sigma_block.weight"):
        out = np.full(shape, 0.7)
    elif key.endswith("sigma_block.bias"):
        out = np.full(shape, 0.2)
    elif key.endswith("sigma_block.freq"):
        out = 4.0 * r.standard_normal(shape)
    elif "bias" in key.rsplit(".", 1)[-1]:
        out = 0.05 * r.standard_normal(shape)
    elif len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        out = r.standard_normal(shape) / np.sqrt(fan_in)
    else:
        out = 0.1 * r.standard_normal(shape)
    return out.astype(np.float32)


def det_state_dict(manifest, seed=0):
    """manifest: {key: shape}.  Returns {key: torch.Tensor} for every non constructor buffer."""
    out = {}
    for key in sorted(manifest):
        if is_constructor_buffer(key):
            continue
        if key.endswith("weight_g"):
            vkey = key[: -len("weight_g")] + "weight_v"
            v = det_tensor(vkey, manifest[vkey], seed)
            arr = det_tensor(key, manifest[key], seed, v_for_g=v)
        else:
            arr = det_tensor(key, manifest[key], seed)
        out[key] = torch.from_numpy(arr.reshape(tuple(manifest[key])))
    return out


def det_lora_factors(named_shapes, seed=0):
    """Deterministic values for the LoRA factors ({key: shape} of the ``lora_*`` parameters): both factors
    non-zero (the reference initialises one of them to zero, i.e. a no-op adapter) and small enough that
    the merged weights stay in the regime of the base weights (delta ~ 10 % of W)."""
    out = {}
    for key in sorted(named_shapes):
        shape = tuple(named_shapes[key])
        r = _rng(seed, key)
        if key.endswith(("lora_weight_a", "lora_linear_a")):          # (out, rank)
            arr = r.standard_normal(shape) * 0.3 / np.sqrt(shape[1])
        else:                                                          # (rank, in [* k])
            arr = r.standard_normal(shape) * 0.3 / np.sqrt(shape[1])
        out[key] = torch.from_numpy(arr.astype(np.float32))
    return out


def det_audio(shape, seed, level=0.05):
    """White-noise 'noisy speech' (synthetic white-noise inputs)."""
    r = np.random.default_rng([seed, 1])
    return torch.from_numpy((level * r.standard_normal(shape)).astype(np.float32))


def det_noise(n, shape, seed):
    """n unit-variance diffusion-noise tensors, in the draw order of universe.py:326,338."""
    r = np.random.default_rng([seed, 2])
    return [torch.from_numpy(r.standard_normal(shape).astype(np.float32)) for _ in range(n)]


def subsample(t, max_n=4096):
    flat = t.detach().reshape(-1)
    stride = max(1, flat.numel() // max_n)
    return flat[::stride].clone(), stride
