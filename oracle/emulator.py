"""PyTorch emulator for the engine's op list -- TEST INFRASTRUCTURE, never on the product path.

Executes an ``open_universe_b200.engine.program.Program`` op by op with ATen on the CPU,
following the kernel contracts of ``include/ou_b200.h``.  Purpose:
  * ``quant=False`` (exact fp32): proves that the host-side lowering (weight-norm /
    anti-alias folding, epilogue wiring, buffer lengths) is algebraically identical to the
    oracle -- runs in the CPU test-suite, no GPU needed;
  * ``quant=True``: additionally rounds weights and every stored activation to ``QDTYPE`` exactly
    where the CUDA kernels do, which bounds the error the precision policy (16-bit storage and
    MMA operands, fp32 accumulation / epilogues / GRU state) can introduce -- the measured
    numbers are quoted in DESIGN.md.
Activations are kept as plain (B, C, T) float tensors here; only the values matter.
"""
import math

import torch
import torch.nn.functional as F

from open_universe_b200.engine import program as P


# storage / MMA-operand type of the emulated kernels when ``quant`` is on: fp16 is the library's
# default policy (csrc/common.cuh); GPU tests set it from ``lib.act_dtype()`` and
# tools/precision_study.py switches it to compare policies (torch.bfloat16 | torch.float16)
QDTYPE = torch.float16


def _q(x, quant):
    return x.to(QDTYPE).float() if quant else x


def prelu(x, a):
    return torch.where(x >= 0, x, a * x)


def run_conv(op, bufs, film=None, quant=False):
    fc = op.fc
    x = bufs[op.src]                                   # (B, Cin, t_in)
    B = x.shape[0]
    assert x.shape[1] == fc.cin and x.shape[2] == op.t_in, (op.name, x.shape, fc.cin, op.t_in)
    if fc.prelu_in is not None:
        x = _q(prelu(x, fc.prelu_in), quant)
    s, taps, rows = fc.s, fc.taps, op.rows
    # gather A[j, q, r*Cin+ci] = x[ci, (j + tap_off + q)*s + r], zero outside [0, t_in)
    lo = fc.tap_off * s
    hi = (rows - 1 + fc.tap_off + taps - 1) * s + s   # exclusive
    pad_l = max(0, -lo)
    pad_r = max(0, hi - op.t_in)
    xp = F.pad(x, (pad_l, pad_r))
    base = lo + pad_l
    seg = xp[:, :, base: base + (rows + taps - 1) * s]          # (B, Cin, (rows+taps-1)*s)
    seg = seg.reshape(B, fc.cin, rows + taps - 1, s)             # [b, ci, jj, r]
    seg = seg.permute(0, 3, 1, 2).reshape(B, s * fc.cin, rows + taps - 1)  # c' = r*Cin+ci
    w = _q(fc.w, quant)                                          # (N, taps, s*Cin)
    acc = F.conv1d(seg, w.permute(0, 2, 1).contiguous())         # (B, N, rows)
    acc = acc + fc.bias[None, :, None]
    if op.dst_kind == "f32_blk":   # logical view (B, rows, N); the device layout is blocked
        bufs[op.dst] = acc.transpose(1, 2).contiguous()          # (B, rows, N)
        return
    # depth-to-space: n = p*Cout + co -> t = j*up + p
    y = acc.reshape(B, fc.up, fc.cout, rows).permute(0, 2, 3, 1).reshape(B, fc.cout, rows * fc.up)
    y = y[:, :, : op.t_out]
    if op.add1 is not None:
        y = y + bufs[op.add1]
    y = y * op.scale1
    if op.add2 is not None:
        y = y + bufs[op.add2]
    y = y * op.scale2
    if op.film_off is not None:
        g = film[:, op.film_off: op.film_off + fc.cout, None]
        b = film[:, op.film_off + fc.cout: op.film_off + 2 * fc.cout, None]
        y = g * y + b
    if op.prelu_out is not None:
        y = prelu(y, op.prelu_out)
    if op.prelu_out2 is not None:
        y = prelu(y, op.prelu_out2)
    bufs[op.dst] = _q(y, quant)


def run_input_conv(op, bufs, in_scale=None, quant=False):
    x = bufs[op.src]                                   # (B, 1, T)
    if op.use_in_scale and in_scale is not None:
        x = x * in_scale[:, None, None]
    y = F.conv1d(x, op.w[:, None, :], op.bias, padding="same")
    bufs[op.dst] = _q(y, quant)


def run_output(op, bufs, x, coef=None, noise=None):
    """Returns (net, x_new)."""
    src = bufs[op.src]
    net = F.conv1d(src, op.w[None], None, padding="same") + op.bias
    net = F.pad(net, (0, op.t_out - net.shape[-1]))
    if coef is None:
        return net, None
    ca, cb, cc = (coef[:, i, None, None] for i in range(3))
    x_new = ca * x + cb * net
    if noise is not None:
        x_new = x_new + cc * noise
    return net, x_new


def _h16(x):
    return x.to(torch.float16).float()


def run_gru(op, bufs, quant=False):
    """quant: the recurrent product W_hh . h_{t-1} takes fp16-rounded operands (fp32 accumulate), as
    ou_gru_bidir's tensor-core kernel does; gates, the state update and h itself stay fp32."""
    gx = bufs[op.src]                                  # (B, T, 6H)
    B, T, _ = gx.shape
    H = op.hidden
    outs = []
    for d in range(2):
        xp = gx[:, :, d * 3 * H: (d + 1) * 3 * H]
        w, b = op.w_hh[d], op.b_hh[d]
        if quant:
            w = _h16(w)
        h = gx.new_zeros(B, H)
        out = gx.new_zeros(B, T, H)
        for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
            hp = (_h16(h) if quant else h) @ w.t() + b
            r = torch.sigmoid(xp[:, t, :H] + hp[:, :H])
            z = torch.sigmoid(xp[:, t, H:2 * H] + hp[:, H:2 * H])
            n = torch.tanh(xp[:, t, 2 * H:] + r * hp[:, 2 * H:])
            h = (1.0 - z) * n + z * h
            out[:, t] = h
        outs.append(out)
    y = torch.cat(outs, dim=-1).transpose(1, 2)        # (B, 2H, T)
    if op.add is not None:
        y = y + bufs[op.add]
    bufs[op.dst] = _q(y * op.scale, quant)


def run_mel(op, bufs, quant=False):
    x = bufs[op.src]                                   # (B, 1, T)
    B = x.shape[0]
    total = (op.frames - 1) * op.hop + op.n_fft
    xp = F.pad(x[:, 0], (op.pad_left, total - op.pad_left - op.t))
    fr = xp.unfold(-1, op.n_fft, op.hop) * op.window   # (B, frames, n_fft)
    n = torch.arange(op.n_fft, dtype=torch.float64)
    k = torch.arange(op.n_fft // 2 + 1, dtype=torch.float64)
    ang = 2.0 * math.pi * torch.outer(n, k) / op.n_fft
    re = fr.double() @ torch.cos(ang)
    im = fr.double() @ torch.sin(ang)
    power = (re.square() + im.square()).float()
    mel = (power @ op.fb).transpose(1, 2)              # (B, n_mels, frames)
    norm = mel.square().sum(dim=1, keepdim=True).mean(dim=-1, keepdim=True).sqrt()
    mel = mel / norm.clamp(min=1e-5)
    bufs[op.dst + ".f32"] = mel
    bufs[op.dst] = _q(mel, quant)


def film_table(prog, g):
    """g: (rows, noise_cond_dim) sigma embedding -> (rows, film_cols) FiLM table."""
    from open_universe_b200.engine import fold
    from open_universe_b200.engine.fold import effective_weight
    cols = []
    for lin, off, cout in prog.film_layers:
        w = effective_weight(lin).float()
        cols.append(g @ w.t() + fold.bias_of(lin))
    return torch.cat(cols, dim=1) if cols else None


def run_program(prog, inputs, film=None, in_scale=None, coef=None, noise=None, quant=False):
    """inputs: {buffer name: (B, C, T) float tensor}.  Returns (bufs, net, x_new)."""
    bufs = dict(inputs)
    net = x_new = None
    for op in P.flat_ops(prog.ops):      # a TrunkOp rounds exactly where its three ConvOps do
        if isinstance(op, P.ConvOp):
            run_conv(op, bufs, film, quant)
        elif isinstance(op, P.InputConvOp):
            run_input_conv(op, bufs, in_scale, quant)
        elif isinstance(op, P.OutputOp):
            net, x_new = run_output(op, bufs, inputs.get("x"), coef, noise)
        elif isinstance(op, P.GruOp):
            run_gru(op, bufs, quant)
        elif isinstance(op, P.MelOp):
            run_mel(op, bufs, quant)
        else:
            raise TypeError(op)
    return bufs, net, x_new
