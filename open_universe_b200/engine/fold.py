"""Load-time weight folding: reference layers -> the canonical fused-conv form of the engine.

Every convolution on the hot path (stride-1 'same' convs, strided down convs, transposed up
convs, the k=stride ``st_convs``, 1x1 projections, GRU input projections) is lowered to ONE
kernel shape, an implicit GEMM over the channel-blocked activation layout:

    acc[j, n] = sum_{q < taps} sum_{c' < s*Cin} W[n, q, c'] * X[ci, (j + tap_off + q) * s + r]
                with c' = r * Cin + ci   (space-to-depth by ``s`` on the input side)
    output channel co = n % Cout, output time t = j * up + n // Cout   (depth-to-space by ``up``)

Folds applied here (all exact in exact arithmetic; SURVEY.md section 7 "verified-exact folds"):
  * weight norm  w = g * v / ||v||  (norm over every dim but 0)           blocks.py:36-41
  * binomial anti-alias low-pass composed into the rate-change conv         blocks.py:205-227
      down:  lowpass(2s+1) -> conv(k=s, stride=s)     ==  conv(k=3s, stride=s, pad=s)
      up:    convT(k=s, stride=s) -> lowpass(2s+1)    ==  convT(k=3s, stride=s, pad=s)
    followed by the separate per-channel bias.
Computed in float64, returned in float32.
"""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class FoldedConv:
    w: torch.Tensor            # (N, taps, s*Cin) fp32
    bias: torch.Tensor         # (N,) fp32
    cin: int
    cout: int
    s: int = 1                 # input space-to-depth factor (stride of a down conv)
    up: int = 1                # output depth-to-space factor (stride of an up conv)
    taps: int = 1
    tap_off: int = 0
    prelu_in: Optional[float] = None

    @property
    def n(self):
        return self.up * self.cout


def effective_weight(mod):
    """Weight of a (possibly old-style weight-normed, possibly LoRA-adapted) Conv1d /
    ConvTranspose1d / Linear.  A LoRA adapter (``open_universe_b200.lora``; reference lora/lora.py:53-56,
    134-137, 213-215) contributes  W + alpha / rank * A B  -- merged here, at load time, so that the
    adapted network runs through exactly the same kernels."""
    if hasattr(mod, "lora_merged_weight"):
        return mod.lora_merged_weight().detach().double()
    if hasattr(mod, "weight_g"):
        v = mod.weight_v.detach().double()
        g = mod.weight_g.detach().double()
        dims = tuple(range(1, v.ndim))
        return v * (g / v.norm(2, dim=dims, keepdim=True))
    return mod.weight.detach().double()


def inner(mod):
    """The Conv1d / ConvTranspose1d / Linear behind a LoRA adapter (or ``mod`` itself)."""
    return mod.lora_inner() if hasattr(mod, "lora_inner") else mod


def bias_of(mod):
    """fp32 bias (or None) of a possibly LoRA-adapted layer."""
    b = getattr(inner(mod), "bias", None)
    return None if b is None else b.detach().float()


def _bias(mod, n, like):
    b = getattr(inner(mod), "bias", None)
    if b is None:
        return torch.zeros(n, dtype=torch.float64, device=like.device)
    return b.detach().double()


def prelu_slope(prelu):
    if prelu is None:
        return None
    w = prelu.weight.detach()
    if w.numel() != 1:
        raise NotImplementedError("per-channel PReLU is not used by the reference (blocks.py:183)")
    return float(w.reshape(-1)[0].item())


def fold_same_conv(conv, prelu=None):
    """Conv1d, stride 1, odd kernel, 'same' zero padding (blocks.py conv1/conv2/conv3, 1x1s)."""
    w = effective_weight(conv)                      # (Cout, Cin, k)
    cout, cin, k = w.shape
    if k % 2 != 1:
        raise NotImplementedError("even 'same' kernels are not on the hot path")
    wf = w.permute(0, 2, 1).contiguous()            # (N, taps, Cin)
    return FoldedConv(wf.float(), _bias(conv, cout, w).float(), cin, cout, 1, 1, k, -(k // 2),
                      prelu_slope(prelu))


def fold_linear_as_conv(weight, bias):
    """(N, Cin) matrix applied per time step (GRU input projection, score.py:83-89)."""
    w = weight.detach().double()
    n, cin = w.shape
    b = torch.zeros(n, dtype=torch.float64, device=w.device) if bias is None else bias.detach().double()
    return FoldedConv(w[:, None, :].contiguous().float(), b.float(), cin, n, 1, 1, 1, 0, None)


def fold_down_conv(pc):
    """PReLU_Conv with Conv1d(k=s, stride=s) (+ binomial low-pass before) -- blocks.py:205-227.
    Also the conditioner's ``st_convs`` (k = stride = 20/80/160, condition.py:33-65)."""
    w = effective_weight(pc.conv)                   # (Cout, Cin, k)
    cout, cin, k = w.shape
    s = pc.stride
    if k != s:
        raise NotImplementedError("rate-change convs have kernel == stride upstream")
    if pc.antialiasing:
        b = pc.low_pass_filter.weights.detach().double()          # (2s+1,)
        # full convolution of the conv taps with the (symmetric) low-pass taps -> 3s taps
        full = torch.zeros(cout, cin, 3 * s, dtype=torch.float64, device=w.device)
        for i in range(s):
            full[:, :, i:i + 2 * s + 1] += w[:, :, i:i + 1] * b
        taps, tap_off = 3, -1
        bias = pc.bias.detach().double() if pc.bias is not None else torch.zeros(
            cout, dtype=torch.float64, device=w.device)
    else:
        full, taps, tap_off = w, 1, 0
        bias = _bias(pc.conv, cout, w)
    # u = q*s + r  ->  W[n, q, r*Cin + ci]
    wf = full.reshape(cout, cin, taps, s).permute(0, 2, 3, 1).reshape(cout, taps, s * cin)
    return FoldedConv(wf.contiguous().float(), bias.float(), cin, cout, s, 1, taps, tap_off,
                      prelu_slope(pc.prelu))


def fold_up_conv(pc):
    """PReLU_Conv with ConvTranspose1d(k=s, stride=s) (+ binomial low-pass after)."""
    w = effective_weight(pc.conv)                   # (Cin, Cout, k)
    cin, cout, k = w.shape
    s = pc.stride
    if k != s:
        raise NotImplementedError("rate-change convs have kernel == stride upstream")
    if pc.antialiasing:
        b = pc.low_pass_filter.weights.detach().double()          # (2s+1,)
        # out[j*s+p] = sum_d sum_ci x[ci, j+d] * sum_i w[ci,co,i] * b[d*s + i + s - p]
        wf = torch.zeros(s, cout, 3, cin, dtype=torch.float64, device=w.device)
        for p in range(s):
            for d in (-1, 0, 1):
                for i in range(s):
                    m = d * s + i + s - p
                    if 0 <= m <= 2 * s:
                        wf[p, :, d + 1, :] += (w[:, :, i] * b[m]).t()
        taps, tap_off = 3, -1
        bias = pc.bias.detach().double() if pc.bias is not None else torch.zeros(
            cout, dtype=torch.float64, device=w.device)
    else:
        wf = w.permute(2, 1, 0)[:, :, None, :]      # (s, Cout, 1, Cin)
        taps, tap_off = 1, 0
        bias = _bias(pc.conv, cout, w)
    wf = wf.reshape(s * cout, taps, cin)            # n = p*Cout + co
    return FoldedConv(wf.contiguous().float(), bias.repeat(s).float(), cin, cout, 1, s, taps,
                      tap_off, prelu_slope(pc.prelu))


def fold_prelu_conv(pc):
    """Dispatch on the PReLU_Conv flavour."""
    if pc.use_transpose:
        return fold_up_conv(pc)
    if pc.stride != 1:
        return fold_down_conv(pc)
    return fold_same_conv(pc.conv, pc.prelu)
