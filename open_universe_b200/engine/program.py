"""Lowering of the UNIVERSE networks to a flat list of fused device ops.

The networks (reference ``score.py:277-297``, ``condition.py:346-377``) are lowered ONCE per
(model weights, batch, length) into a ``Program``: named activation buffers plus a list of
ops, each of which is exactly one CUDA kernel launch of ``csrc/`` through the C ABI
(``include/ou_b200.h``).  The op list is data: ``engine.runtime`` executes it on the GPU;
the test-suite executes the *same* list with a PyTorch emulator to check the host-side
lowering (folding, buffer lengths, epilogue wiring) against the oracle on CPU.

Activation layout in HBM ("blocked"): bf16 ``[B][C/8][T][8]`` -- 8 channels interleaved per
time step so that one time step of one channel group is a 16-byte vector; a tile of
consecutive time steps of one channel group is contiguous (TMA / cp.async friendly) and is
directly the K-major no-swizzle operand layout of the tensor-core MMAs (8 rows x 16 B core
matrices).  Signals (B, 1, T), FiLM tables, GRU pre-activations and GRU state stay fp32.

Fusion plan per ConvBlock (blocks.py:327-412):
    [up conv  : PReLU_in -> convT(+lowpass) + bias -> (+skip)/sqrt2                    ]
    conv1     : PReLU_in -> k5 + bias -> (+sc)/sqrt2 -> FiLM -> PReLU(conv2's)
    conv2     : k3 + bias -> PReLU(conv3's)
    conv3     : k3 + bias -> (+h)/sqrt2 [-> PReLU of the single consumer]
    [down conv: PReLU_in -> (lowpass+)conv stride s + bias                             ]
Step-invariant work hoisted out of the sampler loop: the decoder's ``signal_cond_proj`` 1x1
convs (score.py:165-170,189-194) and the sigma-embedding + FiLM projections for all steps.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import fold
from .fold import FoldedConv

SQRT_HALF = 1.0 / math.sqrt(2.0)


@dataclass
class BufSpec:
    kind: str        # 'blocked' (bf16 [B][C/8][T][8]) | 'f32_blk' (fp32 [B][C/16][T][16]) | 'f32_bt' (fp32 [B][T])
    channels: int
    length: int


@dataclass
class ConvOp:
    """One launch of the implicit-GEMM conv kernel (``ou_conv1d``)."""
    name: str
    src: str
    dst: str
    fc: FoldedConv
    t_in: int
    t_out: int
    rows: int                                 # number of GEMM rows j actually needed
    add1: Optional[str] = None
    scale1: float = 1.0
    add2: Optional[str] = None
    scale2: float = 1.0
    film_off: Optional[int] = None            # column offset of gamma in the FiLM table (beta at +cout)
    prelu_out: Optional[float] = None
    prelu_out2: Optional[float] = None
    dst_kind: str = "blocked"
    flops_exec: float = 0.0                   # FLOPs the kernel executes (folded low-pass included)
    flops_algo: float = 0.0                   # FLOPs of the reference graph for the same layer
    packed: Optional[dict] = None             # device tensors filled in by the runtime


@dataclass
class TrunkOp:
    """conv1 -> FiLM -> conv2 -> conv3 -> residual of one narrow ConvBlock in ONE launch
    (``ou_conv_trunk``, blocks.py:385-399).  ``parts`` are the three ConvOps it stands for: the
    fused kernel rounds to bf16 at exactly their boundaries, so the emulator (and the runtime,
    when the fused kernel is switched off or rejects the shape) simply runs the parts."""
    name: str
    parts: List[ConvOp]
    packed: Optional[dict] = None
    # the next decoder block's transposed up conv (+ skip add) fused behind conv3 (C = 64 only): the block
    # output stays in shared memory, ``parts[2].dst`` is not written by the fused launch
    tail: Optional[ConvOp] = None
    # the network's output conv + EDM / SDE update fused behind conv3 (C = 32 only, ``OutputOp``)
    tail_out: Optional[object] = None
    # the block's own anti-aliased stride-2 down conv fused behind conv3 (C = 32 only): the block output is still
    # written (decoder skip connection) but not read back
    tail_dn: Optional[ConvOp] = None

    def all_parts(self):
        return (self.conv_parts() + ([self.tail_out] if self.tail_out is not None else []))

    def conv_parts(self):
        return (self.parts + ([self.tail] if self.tail is not None else [])
                + ([self.tail_dn] if self.tail_dn is not None else []))

    @property
    def flops_exec(self):
        return sum(p.flops_exec for p in self.conv_parts())

    @property
    def flops_algo(self):
        return sum(p.flops_algo for p in self.conv_parts())


TRUNK_CHANNELS = (32, 64)


def fuse_trunk(prog, name, n_parts=3):
    """Replace the last three ConvOps of ``prog`` (conv1, conv2, conv3 of one block) by a TrunkOp
    when the fused kernel covers them (C in {32, 64}, k5-k3-k3, plain blocked outputs)."""
    c1, c2, c3 = prog.ops[-n_parts:]
    ok = (all(isinstance(o, ConvOp) for o in (c1, c2, c3))
          and c1.fc.cin in TRUNK_CHANNELS
          and all(o.fc.cin == c1.fc.cin and o.fc.cout == c1.fc.cin and o.fc.s == 1 and o.fc.up == 1
                  and o.dst_kind == "blocked" and o.add2 is None and o.t_in == c1.t_in
                  and o.t_out == c1.t_in for o in (c1, c2, c3))
          and (c1.fc.taps, c2.fc.taps, c3.fc.taps) == (5, 3, 3)
          and (c1.fc.tap_off, c2.fc.tap_off, c3.fc.tap_off) == (-2, -1, -1)
          and c1.fc.prelu_in is not None and c1.prelu_out is not None and c1.prelu_out2 is None
          and c2.fc.prelu_in is None and c2.src == c1.dst and c2.add1 is None
          and c2.film_off is None and c2.prelu_out is not None and c2.prelu_out2 is None
          and c3.fc.prelu_in is None and c3.src == c2.dst and c3.add1 == c1.src
          and c3.film_off is None)
    if not ok:
        return False
    del prog.ops[-n_parts:]
    prog.ops.append(TrunkOp(name, [c1, c2, c3]))
    return True


# OU_TRUNK_TAIL=0 keeps the up convs as separate ou_conv1d launches (A/B runs)
import os as _os
TRUNK_TAIL = _os.environ.get("OU_TRUNK_TAIL", "1") != "0"
# OU_TRUNK_OUT_TAIL=0 keeps the output conv + SDE update as a separate ou_output_sde launch (A/B runs)
TRUNK_OUT_TAIL = _os.environ.get("OU_TRUNK_OUT_TAIL", "1") != "0"


def fuse_up_tail(prog):
    """If the last op is a x2 transposed up conv reading the output of the 64-channel TrunkOp right before
    it (decoder: dec.k.trunk -> dec.k+1.up), fold it into that trunk launch as its tail."""
    if not TRUNK_TAIL or len(prog.ops) < 2:
        return False
    up, tr = prog.ops[-1], prog.ops[-2]
    if not (isinstance(up, ConvOp) and isinstance(tr, TrunkOp) and tr.tail is None):
        return False
    c3 = tr.parts[2]
    fc = up.fc
    ok = (c3.fc.cin == 64 and up.src == c3.dst and c3.prelu_out is None and c3.prelu_out2 is None
          and fc.cin == 64 and fc.n == 64 and fc.up == 2 and fc.s == 1
          and (fc.taps, fc.tap_off) in ((3, -1), (1, 0))      # anti-aliased (low-pass folded in) or plain k = s
          and fc.prelu_in is not None and up.add2 is None and up.film_off is None and up.prelu_out is None
          and up.prelu_out2 is None and up.dst_kind == "blocked" and up.t_in == c3.t_out
          and up.t_out <= 2 * up.t_in and up.rows <= up.t_in
          # nothing else may read the block output
          and not any(c3.dst in (getattr(o, "src", None), getattr(o, "add1", None), getattr(o, "add2", None),
                                 getattr(o, "add", None))
                      for o in flat_ops(prog.ops[:-2])))
    if not ok:
        return False
    tr.tail = up
    del prog.ops[-1]
    return True


# OU_TRUNK_DOWN_TAIL=0 keeps enc.0.down as a separate ou_conv1d launch (A/B runs)
TRUNK_DOWN_TAIL = _os.environ.get("OU_TRUNK_DOWN_TAIL", "1") != "0"


def fuse_down_tail(prog):
    """If the last op is the anti-aliased stride-2 down conv (32 -> 64 channels, 3 folded row taps) reading the
    output of the 32-channel TrunkOp right before it (enc.0.trunk -> enc.0.down), fold it into that trunk launch."""
    if not (TRUNK_TAIL and TRUNK_DOWN_TAIL) or len(prog.ops) < 2:
        return False
    dn, tr = prog.ops[-1], prog.ops[-2]
    if not (isinstance(dn, ConvOp) and isinstance(tr, TrunkOp) and tr.tail is None and tr.tail_out is None
            and tr.tail_dn is None):
        return False
    c3, fc = tr.parts[2], dn.fc
    if not (fc.cin == 32 and fc.cout == 64 and fc.s == 2 and fc.up == 1 and (fc.taps, fc.tap_off) in ((3, -1), (1, 0))
            and dn.src == c3.dst and dn.add1 is None and dn.add2 is None and dn.film_off is None
            and dn.prelu_out is None and dn.prelu_out2 is None and dn.dst_kind == "blocked"
            and c3.prelu_out is None and c3.prelu_out2 is None and dn.t_in == c3.t_out
            and dn.t_out == (dn.t_in + 1) // 2):
        return False
    tr.tail_dn = dn
    del prog.ops[-1]
    return True


def fuse_out_tail(prog):
    """If the last op is the network's OutputOp reading the output of the 32-channel TrunkOp right before it
    (dec.last.trunk -> output_conv), fold it into that trunk launch."""
    if not (TRUNK_TAIL and TRUNK_OUT_TAIL) or len(prog.ops) < 2:
        return False
    out, tr = prog.ops[-1], prog.ops[-2]
    if not (isinstance(out, OutputOp) and isinstance(tr, TrunkOp) and tr.tail is None and tr.tail_out is None):
        return False
    c3 = tr.parts[2]
    if not (c3.fc.cin == 32 and out.src == c3.dst and out.w.shape == (32, 3) and out.t == out.t_out == c3.t_out):
        return False
    tr.tail_out = out
    del prog.ops[-1]
    return True


@dataclass
class InputConvOp:
    """(B,1,T) fp32 signal -> blocked bf16 (B,C,T): k-tap 'same' conv of a 1-channel input with an
    optional per-clip input scale (EDM c_in, universe.py:197-203).  ``ou_input_conv``."""
    name: str
    src: str
    dst: str
    w: torch.Tensor                            # (C, k) fp32
    bias: torch.Tensor                         # (C,)
    t: int
    use_in_scale: bool = False
    packed: Optional[dict] = None


@dataclass
class OutputOp:
    """blocked bf16 (B,C,T) -> net(B,1,T) = k-tap conv to ONE channel + bias, fused with the
    EDM mix and the reverse-SDE update (universe.py:197-209, 334-343):
        x_out = ca[b] * x + cb[b] * net + cc[b] * noise      ``ou_output_sde``."""
    name: str
    src: str
    w: torch.Tensor                            # (C, k)
    bias: float
    t: int
    t_out: int                                 # logical signal length (right-padded with zeros)
    packed: Optional[dict] = None


@dataclass
class GruOp:
    """Bidirectional GRU recurrence over pre-computed input projections (``ou_gru_bidir``).
    src: fp32 [B][T][2*3H] (fwd r,z,n | bwd r,z,n, includes b_ih); dst blocked bf16 [B][2H/8][T][8]
    = (h_fwd | h_bwd) optionally (+ add) * scale."""
    name: str
    src: str
    dst: str
    w_hh: torch.Tensor                         # (2, 3H, H) fp32
    b_hh: torch.Tensor                         # (2, 3H)
    hidden: int
    t: int
    add: Optional[str] = None
    scale: float = 1.0
    packed: Optional[dict] = None


@dataclass
class MelOp:
    """(B,1,T) fp32 -> normalised mel 'spectrogram' blocked bf16 (+ fp32 (B,n_mels,F) copy)
    (condition.py:92-108).  ``ou_mel_power`` + ``ou_mel_finalize``."""
    name: str
    src: str
    dst: str
    n_fft: int
    hop: int
    n_mels: int
    pad_left: int
    frames: int
    t: int
    window: torch.Tensor                       # (n_fft,)
    fb: torch.Tensor                           # (n_fft/2+1, n_mels)
    packed: Optional[dict] = None


@dataclass
class Program:
    batch: int
    bufs: Dict[str, BufSpec] = field(default_factory=dict)
    ops: List[object] = field(default_factory=list)
    film_cols: int = 0
    film_layers: List[Tuple[object, int, int]] = field(default_factory=list)  # (linear, offset, cout)
    outputs: Dict[str, str] = field(default_factory=dict)
    meta: dict = field(default_factory=dict)

    def buf(self, name, kind, channels, length):
        spec = BufSpec(kind, channels, length)
        if name in self.bufs and self.bufs[name] != spec:
            raise RuntimeError(f"buffer {name} redefined: {self.bufs[name]} vs {spec}")
        self.bufs[name] = spec
        return name

    def film(self, linear, cout):
        off = self.film_cols
        self.film_layers.append((linear, off, cout))
        self.film_cols += 2 * cout
        return off


def ceil_div(a, b):
    return -(-a // b)


def flat_ops(ops):
    """Op list with every TrunkOp expanded into its three ConvOps."""
    out = []
    for op in ops:
        out.extend(op.all_parts() if isinstance(op, TrunkOp) else [op])
    return out


def add_conv(prog, name, src, dst, fc, t_in, t_out=None, **kw):
    """Append a ConvOp; derives the output length / GEMM row count from the folded geometry."""
    if t_out is None:
        t_out = ceil_div(t_in, fc.s) * fc.up
    rows = ceil_div(t_out, fc.up)
    dst_kind = kw.get("dst_kind", "blocked")
    if dst_kind == "blocked":
        prog.buf(dst, "blocked", fc.cout, t_out)
    else:
        prog.buf(dst, "f32_blk", fc.n, rows)
    op = ConvOp(name, src, dst, fc, t_in, t_out, rows, **kw)
    op.flops_exec = 2.0 * prog.batch * rows * fc.n * fc.taps * fc.s * fc.cin
    if fc.taps == 3 and (fc.s > 1 or fc.up > 1):
        # anti-alias low-pass folded into the rate-change conv: the reference graph runs a k=s
        # conv (1/3 of the folded taps) plus a (2s+1)-tap depthwise FIR (blocks.py:205-227)
        rate = max(fc.s, fc.up)
        fir = fc.cin * t_in if fc.s > 1 else fc.cout * t_out
        op.flops_algo = op.flops_exec / 3.0 + 2.0 * prog.batch * (2 * rate + 1) * fir
    else:
        op.flops_algo = op.flops_exec
    prog.ops.append(op)
    return dst, t_out


def lower_conv_block(prog, blk, pfx, src, t_in, *, film_linear=None, input_cond=None, res=None,
                     length=None, out_prelus=(), raw_cond_out=False, need_tail=True):
    """Lower one ``ConvBlock`` (blocks.py:327-412).  ``src`` holds the RAW block input.
    Returns (out, t_out, skip, cond_out) buffer names (``cond_out`` only if ``raw_cond_out``)."""
    c = blk.n_channels
    h, t = src, t_in
    if blk.rate_change_dir == "up":
        fc = fold.fold_prelu_conv(blk.rate_change_conv)
        t_up = length if length is not None else t_in * fc.up
        if ceil_div(t_up, fc.up) > t_in + 1:
            raise ValueError("target length is more than one frame longer than the upsampled input")
        h, t = add_conv(prog, pfx + ".up", src, pfx + ".h", fc, t_in, t_up,
                        add1=res, scale1=SQRT_HALF if res is not None else 1.0)
        fuse_up_tail(prog)
    elif res is not None:
        raise RuntimeError("lowering expects the residual of a rate-preserving block to be "
                           "pre-added by the producer (GRU epilogue)")
    fc1 = fold.fold_prelu_conv(blk.conv1)
    fc2 = fold.fold_prelu_conv(blk.conv2)
    fc3 = fold.fold_prelu_conv(blk.conv3)
    film_off = prog.film(film_linear, c) if film_linear is not None else None
    cond_out = None
    if raw_cond_out and film_off is None and input_cond is None:
        # the conditioner decoder exports conv1's raw output (condition.py:264-270): keep it raw
        # and let conv2 apply its PReLU on load
        cond_out, _ = add_conv(prog, pfx + ".conv1", h, pfx + ".cond", fc1, t)
        c1 = cond_out
    else:
        if raw_cond_out:
            # module-level ConvBlock.forward with FiLM / conditioning input AND the raw conv1 output
            # requested (blocks.py:385-396): conv1 runs twice, once raw for ``cond_out`` and once with the
            # fused epilogue for conv2 (only the layer-level API takes this path, never the networks)
            cond_out, _ = add_conv(prog, pfx + ".conv1raw", h, pfx + ".cond", fc1, t)
        a2 = fc2.prelu_in
        fc2 = FoldedConv(fc2.w, fc2.bias, fc2.cin, fc2.cout, 1, 1, fc2.taps, fc2.tap_off, None)
        c1, _ = add_conv(prog, pfx + ".conv1", h, pfx + ".c1", fc1, t, add1=input_cond,
                         scale1=SQRT_HALF if input_cond is not None else 1.0, film_off=film_off,
                         prelu_out=a2)
    if not need_tail:
        return None, t, None, cond_out
    a3 = fc3.prelu_in
    fc3 = FoldedConv(fc3.w, fc3.bias, fc3.cin, fc3.cout, 1, 1, fc3.taps, fc3.tap_off, None)
    c2, _ = add_conv(prog, pfx + ".conv2", c1, pfx + ".c2", fc2, t, prelu_out=a3)
    po = list(out_prelus) + [None, None]
    is_down = blk.rate_change_dir == "down"
    v, _ = add_conv(prog, pfx + ".conv3", c2, pfx + ".v", fc3, t, add1=h, scale1=SQRT_HALF,
                    prelu_out=None if is_down else po[0], prelu_out2=None if is_down else po[1])
    if not raw_cond_out:
        fuse_trunk(prog, pfx + ".trunk")
    if is_down:
        fcr = fold.fold_prelu_conv(blk.rate_change_conv)
        out, t_out = add_conv(prog, pfx + ".down", v, pfx + ".out", fcr, t)
        if not raw_cond_out:
            fuse_down_tail(prog)
        return out, t_out, v, cond_out
    return v, t, v, cond_out


def fold_gru_layer(gru, layer):
    """torch.nn.GRU parameters of one bidirectional layer -> (input-projection conv, W_hh, b_hh)."""
    sfx = f"_l{layer}"
    w_ih = torch.cat([getattr(gru, "weight_ih" + sfx), getattr(gru, "weight_ih" + sfx + "_reverse")])
    b_ih = torch.cat([getattr(gru, "bias_ih" + sfx), getattr(gru, "bias_ih" + sfx + "_reverse")])
    w_hh = torch.stack([getattr(gru, "weight_hh" + sfx), getattr(gru, "weight_hh" + sfx + "_reverse")])
    b_hh = torch.stack([getattr(gru, "bias_hh" + sfx), getattr(gru, "bias_hh" + sfx + "_reverse")])
    return (fold.fold_linear_as_conv(w_ih, b_ih), w_hh.detach().float().contiguous(),
            b_hh.detach().float().contiguous())


def lower_gru(prog, gru, pfx, src, t, *, add_last=None, scale_last=1.0):
    """Bidirectional multi-layer GRU over a blocked (B, C, T) buffer (score.py:116, condition.py:213)."""
    h = src
    for layer in range(gru.num_layers):
        fc, w_hh, b_hh = fold_gru_layer(gru, layer)
        gx, _ = add_conv(prog, f"{pfx}.l{layer}.xproj", h, f"{pfx}.gx", fc, t, dst_kind="f32_blk")
        last = layer == gru.num_layers - 1
        dst = prog.buf(f"{pfx}.l{layer}.out", "blocked", 2 * gru.hidden_size, t)
        prog.ops.append(GruOp(f"{pfx}.l{layer}", gx, dst, w_hh, b_hh, gru.hidden_size, t,
                              add=add_last if last else None, scale=scale_last if last else 1.0))
        h = dst
    return h


# OU_HOIST_PRELU=0 leaves the input PReLU of the decoder's up convs to the conv kernel's transform warps (A/B runs)
HOIST_PRELU = _os.environ.get("OU_HOIST_PRELU", "1") != "0"


def hoist_up_prelus(prog):
    """The input PReLU of a transposed up conv (blocks.py:205-227) whose input tensor has no other reader moves
    into the epilogue of the conv that produces that tensor (``prelu_out`` / ``prelu_out2``): the up conv then
    needs no transform pass over its landed tiles, and its four transform warps join the epilogue -- the stage
    that bounds these layers.  Producers inside a fused trunk are left alone (their up conv is the trunk's tail)."""
    if not HOIST_PRELU:
        return 0
    flat = flat_ops(prog.ops)
    in_trunk = {id(c) for op in prog.ops if isinstance(op, TrunkOp) for c in op.all_parts()}
    n = 0
    for up in flat:
        if not (isinstance(up, ConvOp) and up.fc.up > 1 and up.fc.prelu_in is not None and id(up) not in in_trunk):
            continue
        readers = [o for o in flat if up.src in (getattr(o, "src", None), getattr(o, "add1", None),
                                                 getattr(o, "add2", None), getattr(o, "add", None))]
        prod = [o for o in flat if getattr(o, "dst", None) == up.src]
        if len(readers) != 1 or len(prod) != 1 or up.src in getattr(prog, "outputs", {}).values():
            continue
        p = prod[0]
        if not (isinstance(p, ConvOp) and id(p) not in in_trunk and p.dst_kind == "blocked" and p.prelu_out2 is None):
            continue
        if p.prelu_out is None:
            p.prelu_out = up.fc.prelu_in
        else:
            p.prelu_out2 = up.fc.prelu_in
        fc = up.fc
        up.fc = FoldedConv(fc.w, fc.bias, fc.cin, fc.cout, fc.s, fc.up, fc.taps, fc.tap_off, None)
        n += 1
    return n


# --------------------------------------------------------------------------------- score network
def lower_score_network(net, batch, t):
    """ScoreNetwork.forward (score.py:277-297) for a (batch, 1, t) input.

    Program inputs : 'x' (fp32 signal), 'sc{lvl}' (pre-projected conditioning, see
    ``lower_cond_projection``); tables: FiLM (rows x film_cols), in_scale, coefficients.
    Program output : the OutputOp (net output fused with the EDM / SDE update)."""
    if net.precoding is not None:
        raise NotImplementedError("precoding is not used by any shipped config")
    prog = Program(batch)
    c0 = fold.inner(net.input_conv).out_channels
    prog.buf("x", "f32_bt", 1, t)
    w_in = fold.effective_weight(net.input_conv).float()[:, 0, :].contiguous()
    prog.buf("enc.in", "blocked", c0, t)
    prog.ops.append(InputConvOp("input_conv", "x", "enc.in", w_in,
                                fold.bias_of(net.input_conv).contiguous(), t,
                                use_in_scale=True))
    enc, dec = net.encoder, net.decoder
    h, tl = "enc.in", t
    skips, lengths = [], []
    for i, (blk, lin) in enumerate(zip(enc.ds_modules, enc.cond_proj)):
        lengths.append(tl)
        h, tl, skip, _ = lower_conv_block(prog, blk, f"enc.{i}", h, tl, film_linear=lin)
        skips.append(skip)
    if enc.seq_model == "gru":
        if enc.gru_conv_sandwich:
            raise NotImplementedError("encoder_gru_conv_sandwich is false in every shipped config")
        # decoder block 0's "(h + res)/sqrt2" (blocks.py:376) is folded into the GRU epilogue when
        # that block does not change the rate (extra_conv_block=True); otherwise the up conv adds it
        first = dec.up_modules[0]
        if first.rate_change_dir == "none":
            h = lower_gru(prog, enc.gru, "gru", h, tl, add_last=skips[-1], scale_last=SQRT_HALF)
        else:
            h = lower_gru(prog, enc.gru, "gru", h, tl)
    elif enc.seq_model != "none":
        raise ValueError("Values for 'seq_model' can be gru|attention|none")
    skips, lengths = skips[::-1], lengths[::-1]
    n_dec = len(dec.up_modules)
    for lvl, (blk, lin) in enumerate(zip(dec.up_modules, dec.noise_cond_proj)):
        last = lvl == n_dec - 1
        po = ()
        if last:
            po = (fold.prelu_slope(net.prelu), fold.prelu_slope(net.output_conv.prelu))
        cs = blk.n_channels
        prog.buf(f"sc{lvl}", "blocked", cs, lengths[lvl])
        if blk.rate_change_dir == "none":
            if enc.seq_model != "gru":
                raise NotImplementedError("rate-preserving decoder block without GRU")
            h, tl, _, _ = lower_conv_block(prog, blk, f"dec.{lvl}", h, tl, film_linear=lin,
                                           input_cond=f"sc{lvl}", out_prelus=po)
        else:
            h, tl, _, _ = lower_conv_block(prog, blk, f"dec.{lvl}", h, tl, film_linear=lin,
                                           input_cond=f"sc{lvl}", res=skips[lvl],
                                           length=lengths[lvl], out_prelus=po)
    oc = net.output_conv
    if fold.inner(oc.conv).out_channels != 1:
        raise NotImplementedError("output_channels != 1")
    w_out = fold.effective_weight(oc.conv)[0].float().contiguous()      # (C, k)
    b_out = fold.bias_of(oc.conv)
    b_out = float(b_out[0].item()) if b_out is not None else 0.0
    prog.ops.append(OutputOp("output_conv", h, w_out, b_out, tl, t))
    fuse_out_tail(prog)
    hoist_up_prelus(prog)
    prog.meta["lengths"] = lengths
    prog.meta["cond_channels"] = [b.n_channels for b in dec.up_modules]
    return prog


def lower_cond_projection(net, batch, lengths):
    """Step-invariant 1x1 ``signal_cond_proj`` convs (score.py:165-170,189-194,206): run once per
    ``enhance()`` on the conditioner's output instead of once per step.  'cond{lvl}' -> 'sc{lvl}'."""
    prog = Program(batch)
    for lvl, (proj, tl) in enumerate(zip(net.decoder.signal_cond_proj, lengths)):
        fc = fold.fold_same_conv(proj)
        prog.buf(f"cond{lvl}", "blocked", fc.cin, tl)
        add_conv(prog, f"signal_cond_proj.{lvl}", f"cond{lvl}", f"sc{lvl}", fc, tl)
    return prog


# --------------------------------------------------------------------------------- conditioner
def lower_conditioner(net, batch, t, need_signal_tail=True):
    """ConditionerNetwork.forward (condition.py:346-377) for (batch, 1, t) inputs 'x' and 'x_wav'.
    Outputs: 'cond{lvl}' raw conv1 outputs of the decoder blocks, 'h' latent, 'y_hat'."""
    if net.precoding is not None:
        raise NotImplementedError("precoding is not used by any shipped config")
    prog = Program(batch)
    prog.buf("x", "f32_bt", 1, t)
    prog.buf("x_wav", "f32_bt", 1, t)
    mel = net.input_mel
    hop = mel.ds_factor
    frames = ceil_div(t, hop)
    prog.buf("mel", "blocked", mel.n_mels, frames)
    prog.ops.append(MelOp("mel", "x_wav", "mel", mel.n_fft, hop, mel.n_mels, mel.pad_left, frames,
                          t, mel.mel_spec.spectrogram.window.detach().float(),
                          mel.mel_spec.mel_scale.fb.detach().float()))
    fcm = fold.fold_same_conv(mel.conv)
    m0, _ = add_conv(prog, "mel.conv", "mel", "mel.c", fcm, frames)
    x_mel, _, _, _ = lower_conv_block(prog, mel.conv_block, "mel.block", m0, frames)

    c0 = fold.inner(net.input_conv).out_channels
    w_in = fold.effective_weight(net.input_conv)[:, 0, :].float().contiguous()
    prog.buf("enc.in", "blocked", c0, t)
    prog.ops.append(InputConvOp("input_conv", "x", "enc.in", w_in,
                                fold.bias_of(net.input_conv).contiguous(), t))
    enc = net.encoder
    h, tl = "enc.in", t
    lengths = []
    acc = x_mel
    n_sum = 1
    st_outs = []
    for i, blk in enumerate(enc.ds_modules):
        lengths.append(tl)
        h, tl_next, skip, _ = lower_conv_block(prog, blk, f"enc.{i}", h, tl)
        st = enc.st_convs[i] if i < len(enc.st_convs) else None
        if st is not None:
            fcs = fold.fold_down_conv(st)
            acc, t_st = add_conv(prog, f"enc.st{i}", skip, f"enc.st{i}.sum", fcs, tl, add1=acc,
                                 scale1=1.0)
            if t_st != frames:
                raise ValueError("st_conv output length does not match the mel frame count")
            n_sum += 1
            st_outs.append(acc)
        tl = tl_next
    # out = (x_mel + sum(st_convs) + x) / sqrt(n+1)   (condition.py:202-206): the running sum is
    # carried through the st_conv epilogues; the last op producing ``h`` adds it and scales.
    n_sum += 1
    last = next((op for op in reversed(flat_ops(prog.ops)) if isinstance(op, ConvOp) and op.dst == h),
                None)
    if last is None or last.add2 is not None:
        raise RuntimeError("unexpected producer of the encoder output")
    if any(isinstance(op, TrunkOp) and last in op.all_parts() for op in prog.ops):
        # the fused trunk kernel (and its tails) has no second residual input: run this block conv by conv
        idx = next(i for i, op in enumerate(prog.ops) if isinstance(op, TrunkOp) and last in op.all_parts())
        prog.ops[idx:idx + 1] = prog.ops[idx].all_parts()
    if tl != frames:
        raise ValueError("encoder output length does not match the mel frame count")
    last.add2, last.scale2 = acc, 1.0 / math.sqrt(n_sum)
    if enc.seq_model != "gru":
        raise ValueError("Values for 'seq_model' can be gru|attention")
    out1, _, _, _ = lower_conv_block(prog, enc.conv_block1, "enc.cb1", h, tl)
    g = lower_gru(prog, enc.gru, "gru", out1, tl,
                  add_last=out1 if enc.with_gru_residual else None,
                  scale_last=SQRT_HALF if enc.with_gru_residual else 1.0)
    lat, _, _, _ = lower_conv_block(prog, enc.conv_block2, "enc.cb2", g, tl)
    prog.outputs["h"] = lat
    lengths = lengths[::-1]
    dec = net.decoder
    y, _, _, _ = lower_conv_block(prog, dec.input_conv_block, "dec.in", lat, tl)
    n_up = len(dec.up_modules)
    for lvl, (blk, length) in enumerate(zip(dec.up_modules, lengths)):
        last_blk = lvl == n_up - 1
        y, tl, _, cond = lower_conv_block(prog, blk, f"dec.{lvl}", y, tl, length=length,
                                          raw_cond_out=True,
                                          need_tail=need_signal_tail or not last_blk)
        prog.outputs[f"cond{lvl}"] = cond
    if need_signal_tail:
        if net.output_conv is not None:
            fco = fold.fold_same_conv(net.output_conv)
            y, _ = add_conv(prog, "output_conv", y, "y_hat", fco, tl)
        prog.outputs["y_hat"] = y
    prog.meta["lengths"] = lengths
    prog.meta["t_final"] = tl
    return prog


# --------------------------------------------------------------------------------- work model
def op_bytes(op, batch, elem=2):
    """Algorithmic HBM bytes of ONE launch of ``op`` in this design: every input tensor read once, the
    output written once, weights once (``elem`` = bytes per stored activation).  The roofline numerators
    of bench.py / tools/profile_layers.py; ncu's dram__bytes are compared against them."""
    if isinstance(op, TrunkOp):
        c1 = op.parts[0]
        c, t = c1.fc.cin, c1.t_in
        w = sum(p.fc.w.numel() for p in op.conv_parts()) * elem
        if op.tail_out is not None:  # x (+ sc) in; signal x, noise in and x out (fp32); the block output stays on chip
            return elem * batch * c * t * (1 + (c1.add1 is not None)) + 3 * 4 * batch * t + w
        if op.tail_dn is not None:   # x (+ sc) in; block output (skip connection) and the down conv's output out
            dn = op.tail_dn
            return (elem * batch * c * t * (2 + (c1.add1 is not None)) + elem * batch * dn.fc.cout * dn.t_out + w)
        if op.tail is not None:      # x (+ sc) in; the up conv's skip in and output out; the block output stays on chip
            up = op.tail
            return (elem * batch * c * t * (1 + (c1.add1 is not None))
                    + elem * batch * up.fc.cout * up.t_out * (1 + (up.add1 is not None)) + w)
        return elem * batch * c * t * (2 + (c1.add1 is not None)) + w
    if isinstance(op, ConvOp):
        fc = op.fc
        n_in = elem * batch * fc.cin * op.t_in
        w = fc.w.numel() * elem
        if op.dst_kind != "blocked":
            return n_in + 4 * batch * op.rows * fc.n + w
        n_add = (op.add1 is not None) + (op.add2 is not None)
        return n_in + elem * batch * fc.cout * op.t_out * (1 + n_add) + w
    if isinstance(op, InputConvOp):
        return 4 * batch * op.t + elem * batch * op.w.shape[0] * op.t
    if isinstance(op, OutputOp):
        return elem * batch * op.w.shape[0] * op.t + 3 * 4 * batch * op.t_out
    if isinstance(op, GruOp):
        h = op.hidden
        return (4 * batch * op.t * 6 * h + elem * batch * op.t * 2 * h * (1 + (op.add is not None))
                + 4 * op.w_hh.numel())
    if isinstance(op, MelOp):
        return 4 * batch * op.t + (4 + elem) * batch * op.n_mels * op.frames
    raise TypeError(op)


def op_flops(op, batch):
    """Algorithmic FLOPs of one launch (reference graph as written, see ``add_conv``)."""
    if isinstance(op, TrunkOp):
        return op.flops_algo + (op_flops(op.tail_out, batch) if op.tail_out is not None else 0.0)
    if isinstance(op, ConvOp):
        return op.flops_algo
    if isinstance(op, (InputConvOp, OutputOp)):
        return 2.0 * batch * op.t * op.w.numel()
    if isinstance(op, GruOp):
        return 2.0 * batch * op.t * op.w_hh.numel()
    if isinstance(op, MelOp):
        return 2.0 * batch * op.frames * (op.n_fft * (op.n_fft + 2) + (op.n_fft // 2 + 1) * op.n_mels)
    raise TypeError(op)


def kernel_name(op):
    """Kernel family an op launches (bench.py's per-kernel breakdown)."""
    if isinstance(op, TrunkOp):
        return f"trunk_kernel<{op.parts[0].fc.cin}>"
    return {ConvOp: "conv1d_tc_kernel", InputConvOp: "input_conv_kernel", OutputOp: "output_sde_kernel",
            GruOp: "gru_cluster_f16_kernel", MelOp: "mel_power+finalize"}[type(op)]
