"""GPU executor: runs ``engine.program`` op lists through the C ABI of ``libou_b200.so``.

PyTorch is used for device memory (``torch.empty`` / ``Tensor.data_ptr``), the current CUDA
stream and tiny host-side scalar math only; every FLOP of the networks is issued by a kernel of
``csrc/``.  There is NO CPU path: any entry point called with non-CUDA tensors raises.
"""
import functools
import math
import os
from collections import OrderedDict
from ctypes import byref, c_void_p

import torch

from . import fold, lib
from . import program as P


class NoCudaPathError(RuntimeError):
    pass


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NoCudaPathError(
                "open_universe_b200 computes on CUDA devices only (sm_100a kernels through "
                "libou_b200.so); there is no CPU fallback -- move the model and inputs to 'cuda'")


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _first_cuda_tensor(obj):
    if torch.is_tensor(obj):
        return obj if obj.is_cuda else None
    if isinstance(obj, (list, tuple)):
        for o in obj:
            t = _first_cuda_tensor(o)
            if t is not None:
                return t
    return None


def on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its first CUDA tensor argument made current: kernels are
    launched through ctypes on ``torch.cuda.current_stream()``, which belongs to the CURRENT device,
    not to the device the tensors live on (``model.to('cuda:1')`` without ``torch.cuda.set_device``)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(list(args) + list(kwargs.values()))
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def same_device(*tensors):
    """All CUDA tensors of one call must live on one device (pointers are passed raw to the kernels)."""
    dev = None
    for t in tensors:
        if t is None or not torch.is_tensor(t):
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices in one call: {dev} and {t.device}")
    return dev


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def round_up(a, b):
    return -(-a // b) * b


# ------------------------------------------------------------------------------------ packing
def choose_npad(n):
    if n >= 128:
        return round_up(n, 128)
    if n > 32:
        return round_up(n, 64)
    return 32


def channel_block(c):
    """Channel block CB of the blocked activation layout [B][C/CB][T][CB] (csrc/common.cuh cl_cb)."""
    if c % 16:
        raise ValueError(f"blocked tensors need a multiple of 16 channels, got {c}")
    return 64 if c % 64 == 0 else (32 if c % 32 == 0 else 16)


def alloc_blocked(b, c, t, device):
    cb = channel_block(c)
    return torch.empty(b, c // cb, t, cb, dtype=lib.act_dtype(), device=device)


def pack_conv_weights(fc, device):
    """(N, taps, K) fp32 -> bf16 [taps][kpad/8][npad][8] (ou_conv_params.w, mma.sync / naive kernels)
    and, for stride-1 geometry, [taps][cin/CB][npad][CB] K-major tiles (ou_conv_params.w_tc, tcgen05)."""
    n, taps, k = fc.w.shape
    kpad, npad = round_up(k, 32), choose_npad(n)
    src = fc.w.to(device).permute(1, 2, 0)                       # (taps, K, N)
    w = torch.zeros(taps, kpad, npad, dtype=torch.float32, device=device)
    w[:, :k, :n] = src
    w = w.reshape(taps, kpad // 8, 8, npad).permute(0, 1, 3, 2).contiguous()
    # K blocks of the tcgen05 path: c' = r*cin + ci split in channel blocks of the input layout
    cb = channel_block(fc.cin)
    wt = torch.zeros(taps, k, npad, dtype=torch.float32, device=device)
    wt[:, :, :n] = src
    wt = wt.reshape(taps, k // cb, cb, npad).permute(0, 1, 3, 2).contiguous()
    dt = lib.act_dtype()
    return {"w": w.to(dt), "bias": fc.bias.to(device).contiguous(), "kpad": kpad,
            "npad": npad, "w_tc": wt.to(dt)}


def alloc_buffers(prog, device, skip=()):
    bufs = {}
    b = prog.batch
    for name, spec in prog.bufs.items():
        if name in skip:
            continue
        if spec.kind == "blocked":
            bufs[name] = alloc_blocked(b, spec.channels, spec.length, device)
        elif spec.kind == "f32_blk":     # fp32 [B][C/16][T][16] (GRU input pre-activations)
            bufs[name] = torch.empty(b, spec.channels // 16, spec.length, 16, dtype=torch.float32,
                                     device=device)
        elif spec.kind == "f32_bt":
            bufs[name] = None   # supplied by the caller
        else:
            raise ValueError(spec.kind)
    return bufs


def dft_matrix(n_fft):
    """fp32 [n_fft][2 * (n_fft/2 + 1)]: columns (2k, 2k + 1) = (cos, sin)(2 pi i k / n_fft), evaluated in
    double precision on the exact residue i*k mod n_fft (ou_mel_power's GEMM operand)."""
    i = torch.arange(n_fft, dtype=torch.int64)[:, None]
    k = torch.arange(n_fft // 2 + 1, dtype=torch.int64)[None, :]
    ang = 2.0 * math.pi * ((i * k) % n_fft).to(torch.float64) / n_fft
    return torch.stack([torch.cos(ang), torch.sin(ang)], dim=2).reshape(n_fft, -1).float().contiguous()


def prepare_ops(prog, device, shared=None, tag=""):
    """Upload packed weights / small tables for every op of a program.  ``shared``: dict that keeps
    ONE packed copy per (program kind, op name, device) for all runners of a module -- the packed
    weights depend on the weights only, not on (batch, length)."""
    def packed_for(op, make):
        key = (tag, op.name, str(device))
        if shared is not None and key in shared:
            return shared[key]
        pk = make()
        if shared is not None:
            shared[key] = pk
        return pk

    for op in P.flat_ops(prog.ops):
        if isinstance(op, P.ConvOp):
            op.packed = packed_for(op, lambda: pack_conv_weights(op.fc, device))
        elif isinstance(op, P.InputConvOp):
            op.packed = packed_for(op, lambda: {"w": op.w.to(device).contiguous(),
                                                "bias": op.bias.to(device).contiguous()})
        elif isinstance(op, P.OutputOp):
            op.packed = packed_for(op, lambda: {"w": op.w.to(device).contiguous(),
                                                "w_kc": op.w.t().contiguous().to(device)})   # tap-major [k][C]
        elif isinstance(op, P.GruOp):
            op.packed = packed_for(op, lambda: {"w_hh": op.w_hh.to(device).contiguous(),
                                                "b_hh": op.b_hh.to(device).contiguous()})
        elif isinstance(op, P.MelOp):
            tables = packed_for(op, lambda: {"window": op.window.to(device).contiguous(),
                                             "fb": op.fb.to(device).contiguous(),
                                             "dft": dft_matrix(op.n_fft).to(device)})
            op.packed = dict(tables,
                             power=torch.empty(prog.batch * op.frames, op.n_fft // 2 + 1,
                                               dtype=torch.float32, device=device),
                             mel=torch.empty(prog.batch, op.n_mels, op.frames, dtype=torch.float32,
                                             device=device),
                             energy=torch.empty(prog.batch, op.frames, dtype=torch.float32,
                                                device=device))


# ------------------------------------------------------------------------------------ op launch
# bench.py sets this to a list to time every op launch of the programs with CUDA events on the
# launching stream: entries (op, batch, start event, end event)
PROFILE = None


def conv_params(op, bufs, batch, gamma=None, beta=None, film_bstride=0, max_ctas=0):
    """``ou_conv_params`` of a ConvOp bound to device buffers."""
    fc, pk = op.fc, op.packed
    prm = lib.ConvParams()
    prm.max_ctas = max_ctas
    prm.x = bufs[op.src].data_ptr()
    prm.w = pk["w"].data_ptr()
    prm.w_tc = pk["w_tc"].data_ptr() if pk["w_tc"] is not None else None
    prm.bias = pk["bias"].data_ptr()
    prm.add1 = bufs[op.add1].data_ptr() if op.add1 else None
    prm.add2 = bufs[op.add2].data_ptr() if op.add2 else None
    prm.gamma = gamma
    prm.beta = beta
    if op.dst_kind == "blocked":
        prm.out, prm.out_f32_blk = bufs[op.dst].data_ptr(), None
    else:
        prm.out, prm.out_f32_blk = None, bufs[op.dst].data_ptr()
    prm.batch, prm.cin, prm.t_in = batch, fc.cin, op.t_in
    prm.s, prm.taps, prm.tap_off = fc.s, fc.taps, fc.tap_off
    prm.n, prm.cout, prm.up = fc.n, fc.cout, fc.up
    prm.kpad, prm.npad = pk["kpad"], pk["npad"]
    prm.rows, prm.t_out = op.rows, op.t_out
    prm.film_bstride = film_bstride
    prm.has_prelu_in = fc.prelu_in is not None
    prm.prelu_in = fc.prelu_in or 0.0
    prm.has_prelu_out = op.prelu_out is not None
    prm.prelu_out = op.prelu_out or 0.0
    prm.has_prelu_out2 = op.prelu_out2 is not None
    prm.prelu_out2 = op.prelu_out2 or 0.0
    prm.scale1, prm.scale2 = op.scale1, op.scale2
    return prm


def launch_conv(op, bufs, batch, gamma=None, beta=None, film_bstride=0, naive=False, max_ctas=0):
    prm = conv_params(op, bufs, batch, gamma, beta, film_bstride, max_ctas)
    fn = lib.load().ou_conv1d_naive if naive else lib.load().ou_conv1d
    lib.check(fn(byref(prm), _stream()))


# OU_TRUNK=0 runs every fused ConvBlock trunk as its three ou_conv1d launches (A/B measurements)
USE_TRUNK = os.environ.get("OU_TRUNK", "1") != "0"


def trunk_params(op, bufs, batch, film=None, film_bstride=0, max_ctas=0, coef=None, noise=None, xout=None,
                 net_out=None):
    """``ou_trunk_params`` of a TrunkOp bound to device buffers (``coef`` ... ``net_out``: the per-evaluation
    arguments of a fused output tail, as for the OutputOp)."""
    c1, c2, c3 = op.parts
    prm = lib.TrunkParams()
    prm.max_ctas = max_ctas
    prm.x = bufs[c1.src].data_ptr()
    prm.w1, prm.w2, prm.w3 = (c.packed["w_tc"].data_ptr() for c in op.parts)
    prm.b1, prm.b2, prm.b3 = (c.packed["bias"].data_ptr() for c in op.parts)
    prm.sc = bufs[c1.add1].data_ptr() if c1.add1 else None
    prm.gamma = prm.beta = None
    if c1.film_off is not None and film is not None:
        base = film.data_ptr() + 4 * c1.film_off
        prm.gamma, prm.beta = base, base + 4 * c1.fc.cout
    prm.film_bstride = film_bstride
    prm.out = bufs[c3.dst].data_ptr()
    prm.batch, prm.channels, prm.t = batch, c1.fc.cin, c1.t_in
    prm.taps1, prm.taps2, prm.taps3 = c1.fc.taps, c2.fc.taps, c3.fc.taps
    prm.prelu_in, prm.prelu_mid1, prm.prelu_mid2 = c1.fc.prelu_in, c1.prelu_out, c2.prelu_out
    prm.has_prelu_out = c3.prelu_out is not None
    prm.prelu_out = c3.prelu_out or 0.0
    prm.has_prelu_out2 = c3.prelu_out2 is not None
    prm.prelu_out2 = c3.prelu_out2 or 0.0
    prm.scale1, prm.scale3 = c1.scale1, c3.scale1
    if op.tail is not None:
        up = op.tail
        prm.up_w, prm.up_bias = _tail_frame(up).data_ptr(), up.packed["bias"].data_ptr()
        prm.up_taps = up.fc.taps
        prm.up_skip = bufs[up.add1].data_ptr() if up.add1 else None
        prm.up_out = bufs[up.dst].data_ptr()
        prm.up_t_out, prm.up_scale, prm.up_prelu_in = up.t_out, up.scale1, up.fc.prelu_in
    if op.tail_dn is not None:
        dn = op.tail_dn
        if dn.packed["npad"] != dn.fc.cout:
            raise ValueError("down tail: padded output width")
        # w_tc of the 32 -> 64, s = 2, 3-tap conv is [3 taps][2 phases][64][32] = sample tap m = 2 q + r major
        prm.dn_w, prm.dn_bias = _tail_frame(dn).data_ptr(), dn.packed["bias"].data_ptr()
        prm.dn_taps = dn.fc.taps
        prm.dn_out, prm.dn_t_out = bufs[dn.dst].data_ptr(), dn.t_out
        prm.dn_prelu_in = 1.0 if dn.fc.prelu_in is None else dn.fc.prelu_in
    if op.tail_out is not None:
        oc = op.tail_out
        prm.out_w, prm.out_bias = oc.packed["w_kc"].data_ptr(), oc.bias
        prm.out_coef, prm.out_x = _ptr(coef), _ptr(bufs.get("x"))
        prm.out_noise, prm.out_xout, prm.out_net = _ptr(noise), _ptr(xout), _ptr(net_out)
    return prm


def _tail_frame(conv):
    """Tensor-core weight tiles of a trunk tail in the kernel's 3-row-tap frame: a plain k = s conv (1 tap) sits
    in the middle tap, the outer taps are zero (and skipped by the kernel)."""
    pk = conv.packed
    if conv.fc.taps == 3:
        return pk["w_tc"]
    if "w_tc3" not in pk:
        w = pk["w_tc"]
        z = torch.zeros_like(w)
        pk["w_tc3"] = torch.cat([z, w, z], dim=0).contiguous()
    return pk["w_tc3"]


def launch_trunk(op, bufs, batch, film=None, film_bstride=0, max_ctas=0, coef=None, noise=None, xout=None,
                 net_out=None):
    """One ``ou_conv_trunk`` launch for the three ConvOps of a TrunkOp (+ its fused tails)."""
    prm = trunk_params(op, bufs, batch, film, film_bstride, max_ctas, coef, noise, xout, net_out)
    lib.check(lib.load().ou_conv_trunk(byref(prm), _stream()))


def record_plan(exe):
    """``ou_plan`` (include/ou_b200.h, plan level) of an Executor's program: every op with its device pointers
    resolved, replayed by ``ou_plan_run``.  Returns the handle (caller destroys it with ``ou_plan_destroy``)."""
    L = lib.load()
    bufs, B = exe.bufs, exe.batch
    handle = c_void_p()
    lib.check(L.ou_plan_create(byref(handle)))
    for op in exe.prog.ops:
        if isinstance(op, P.TrunkOp):
            off = op.parts[0].film_off
            prm = trunk_params(op, bufs, B, max_ctas=exe.max_ctas)
            lib.check(L.ou_plan_add_trunk(handle, byref(prm), -1 if off is None else off))
        elif isinstance(op, P.ConvOp):
            prm = conv_params(op, bufs, B, max_ctas=exe.max_ctas)
            lib.check(L.ou_plan_add_conv(handle, byref(prm), -1 if op.film_off is None else op.film_off))
        elif isinstance(op, P.InputConvOp):
            c, k = op.packed["w"].shape
            lib.check(L.ou_plan_add_input_conv(handle, _ptr(op.packed["w"]), _ptr(op.packed["bias"]),
                                               _ptr(bufs[op.dst]), B, op.t, c, k, 1 if op.use_in_scale else 0))
        elif isinstance(op, P.OutputOp):
            c, k = op.packed["w"].shape
            lib.check(L.ou_plan_add_output_sde(handle, _ptr(bufs[op.src]), _ptr(op.packed["w"]), op.bias, B,
                                               c, k, op.t, op.t_out))
        elif isinstance(op, P.GruOp):
            lib.check(L.ou_plan_add_gru(handle, _ptr(bufs[op.src]), _ptr(op.packed["w_hh"]),
                                        _ptr(op.packed["b_hh"]), _ptr(bufs[op.add] if op.add else None),
                                        op.scale, _ptr(bufs[op.dst]), B, op.t, op.hidden, exe.gru_cluster))
        elif isinstance(op, P.MelOp):
            pk = op.packed
            lib.check(L.ou_plan_add_mel(handle, _ptr(pk["window"]), _ptr(pk["fb"]), _ptr(pk["dft"]), _ptr(pk["power"]),
                                        _ptr(pk["mel"]), _ptr(pk["energy"]), _ptr(bufs[op.dst]), B, op.t, op.n_fft,
                                        op.hop, op.n_mels, op.pad_left, op.frames))
        else:
            raise TypeError(op)
    assert L.ou_plan_size(handle) == len(exe.prog.ops)
    return handle


class Executor:
    """A lowered program bound to device buffers."""

    def __init__(self, prog, device, external=(), shared=None, tag=""):
        self.prog = prog
        self.device = device
        self.batch = prog.batch
        prepare_ops(prog, device, shared, tag)
        self.bufs = alloc_buffers(prog, device, skip=external)
        self.naive = False   # tests: route ConvOps through the fp32 CUDA-core reference kernel
        self.max_ctas = 0    # cap on the persistent conv grids (0 = all SMs), see PipelinedScoreRunner
        self.gru_cluster = 0  # CTAs per GRU cluster (0 = lowest latency; 4 = half the SMs), see PipelinedScoreRunner

    def run(self, film=None, film_bstride=0, in_scale=None, coef=None, noise=None, xout=None,
            net_out=None, ops=None):
        """Launch the program's ops in order (``ops``: a contiguous part of ``self.prog.ops``)."""
        L = lib.load()
        bufs, B = self.bufs, self.batch
        ops = self.prog.ops if ops is None else ops
        if self.naive or not USE_TRUNK:
            ops = P.flat_ops(ops)
        for op in ops:
            if PROFILE is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            if isinstance(op, P.TrunkOp):
                launch_trunk(op, bufs, B, film, film_bstride, self.max_ctas, coef, noise, xout, net_out)
            elif isinstance(op, P.ConvOp):
                gamma = beta = None
                if op.film_off is not None:
                    base = film.data_ptr() + 4 * op.film_off
                    gamma, beta = base, base + 4 * op.fc.cout
                launch_conv(op, bufs, B, gamma, beta, film_bstride, self.naive, self.max_ctas)
            elif isinstance(op, P.InputConvOp):
                c, k = op.packed["w"].shape
                lib.check(L.ou_input_conv(_ptr(bufs[op.src]), _ptr(op.packed["w"]),
                                          _ptr(op.packed["bias"]),
                                          _ptr(in_scale if op.use_in_scale else None),
                                          _ptr(bufs[op.dst]), B, op.t, c, k, _stream()))
            elif isinstance(op, P.OutputOp):
                c, k = op.packed["w"].shape
                lib.check(L.ou_output_sde(_ptr(bufs[op.src]), _ptr(op.packed["w"]), op.bias,
                                          _ptr(coef), _ptr(bufs["x"]), _ptr(noise), _ptr(xout),
                                          _ptr(net_out), B, c, k, op.t, op.t_out, _stream()))
            elif isinstance(op, P.GruOp):
                lib.check(L.ou_gru_bidir_ex(_ptr(bufs[op.src]), _ptr(op.packed["w_hh"]),
                                            _ptr(op.packed["b_hh"]),
                                            _ptr(bufs[op.add] if op.add else None), op.scale,
                                            _ptr(bufs[op.dst]), B, op.t, op.hidden, self.gru_cluster, _stream()))
            elif isinstance(op, P.MelOp):
                pk = op.packed
                lib.check(L.ou_mel_power(_ptr(bufs[op.src]), _ptr(pk["window"]), _ptr(pk["fb"]),
                                         _ptr(pk["dft"]), _ptr(pk["power"]), _ptr(pk["mel"]),
                                         _ptr(pk["energy"]),
                                         B, op.t, op.n_fft, op.hop, op.n_mels, op.pad_left,
                                         op.frames, _stream()))
                lib.check(L.ou_mel_finalize(_ptr(pk["mel"]), _ptr(pk["energy"]), _ptr(pk["mel"]),
                                            _ptr(bufs[op.dst]), B, op.n_mels, op.frames, _stream()))
            else:
                raise TypeError(op)
            if PROFILE is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                PROFILE.append((op, B, e0, e1))


# ------------------------------------------------------------------------------------ layouts
@on_tensor_device
def pack_blocked(x):
    """(B, C, T) fp32 -> blocked 16-bit [B][C/CB][T][CB] (fp16 / bf16 per ``lib.act_dtype()``)."""
    require_cuda(x)
    b, c, t = x.shape
    x = x.contiguous().float()
    out = alloc_blocked(b, c, t, x.device)
    lib.check(lib.load().ou_pack_blocked(_ptr(x), _ptr(out), b, c, t, _stream()))
    return out


@on_tensor_device
def unpack_blocked(xb):
    """blocked 16-bit [B][C/CB][T][CB] -> (B, C, T) fp32."""
    require_cuda(xb)
    b, nblk, t, cb = xb.shape
    out = torch.empty(b, nblk * cb, t, dtype=torch.float32, device=xb.device)
    lib.check(lib.load().ou_unpack_blocked(_ptr(xb), _ptr(out), b, nblk * cb, t, _stream()))
    return out


def pack_f32_blocked(x):
    """(B, T, N) fp32 -> the blocked fp32 layout [B][N/16][T][16] of the GRU input pre-activations."""
    b, t, n = x.shape
    return x.float().reshape(b, t, n // 16, 16).permute(0, 2, 1, 3).contiguous()


def unpack_f32_blocked(xb):
    """blocked fp32 [B][N/16][T][16] -> (B, T, N)."""
    b, nblk, t, cb = xb.shape
    return xb.permute(0, 2, 1, 3).reshape(b, t, nblk * cb)


# ------------------------------------------------------------------------------------ caching
def weights_version(module):
    """Cheap fingerprint that changes whenever a parameter / buffer is written in place, replaced,
    or moved (EMA swap in eval()/train(), load_state_dict, .to()).  Writes
    through ``p.data`` do NOT bump ``_version``: ``utils.ema`` therefore copies through the parameter
    itself, and ``invalidate(module)`` is the explicit hook for anything else."""
    v = module.__dict__.get("_ou_generation", 0)
    for t in list(module.parameters()) + list(module.buffers()):
        v = (v * 1000003 + t._version + (t.data_ptr() & 0xFFFFFFFF)) & 0xFFFFFFFFFFFF
    return v


def invalidate(module):
    """Drop every cached runner / packed weight of ``module`` (and bump its generation)."""
    for m in module.modules():
        m.__dict__["_ou_generation"] = m.__dict__.get("_ou_generation", 0) + 1
        m.__dict__.pop("_ou_cache", None)


# Per-module cache: runners (lowered program + activation buffers + captured graphs) per
# (kind, batch, length, device) in a small LRU -- a folder of files with many different lengths
# must not accumulate gigabytes of buffers -- and ONE copy of the packed weights shared by all of them.
MAX_CACHED_SHAPES = int(os.environ.get("OU_CACHE_SHAPES", "4"))
MAX_CACHED_LOOPS = 2


def _cache(module):
    c = module.__dict__.get("_ou_cache")
    ver = weights_version(module)
    if c is None or c["version"] != ver:
        c = {"version": ver, "runners": OrderedDict(), "packed": {}}
        module.__dict__["_ou_cache"] = c
    return c


def _get_runner(module, key, make):
    c = _cache(module)
    runners = c["runners"]
    if key in runners:
        runners.move_to_end(key)
        return runners[key]
    r = make(c["packed"])
    runners[key] = r
    # one slot per kind and shape: a score runner and a conditioner runner of the same shape are two entries
    while len(runners) > 2 * MAX_CACHED_SHAPES:
        runners.popitem(last=False)
    return r


# ------------------------------------------------------------------------------------ sigma embedding
def embedding_spec(sb, device):
    """Device-resident parameters of a SigmaBlock / SimpleTimeEmbedding (sigma_block.py:36-78)."""
    if hasattr(sb, "freq"):
        return ("rff", sb.freq.detach().float().to(device).contiguous(), [
            (fold.effective_weight(lyr.lin).float().to(device).contiguous(),
             fold.bias_of(lyr.lin).to(device).contiguous(), fold.prelu_slope(lyr.prelu))
            for lyr in (sb.layer1, sb.layer2, sb.layer3)])
    return ("simple", float(sb.weight.detach().reshape(-1)[0].item()),
            float(sb.bias.detach().reshape(-1)[0].item()))


def sigma_embedding(spec, dim, log10_sigma):
    L = lib.load()
    ls = log10_sigma.contiguous().float()
    device = ls.device
    rows = ls.numel()
    if spec[0] == "simple":
        g = torch.empty(rows, dim, dtype=torch.float32, device=device)
        lib.check(L.ou_sigma_embed_simple(_ptr(ls), spec[1], spec[2], _ptr(g), rows, dim // 2,
                                          _stream()))
        return g
    _, freq, layers = spec
    cur = torch.empty(rows, 2 * freq.numel(), dtype=torch.float32, device=device)
    lib.check(L.ou_sigma_embed_rff(_ptr(ls), _ptr(freq), _ptr(cur), rows, freq.numel(), _stream()))
    for w, b, slope in layers:
        nxt = torch.empty(rows, w.shape[0], dtype=torch.float32, device=device)
        lib.check(L.ou_linear_f32(_ptr(cur), _ptr(w), _ptr(b), _ptr(nxt), rows, w.shape[1],
                                  w.shape[0], w.shape[0], 1, slope, _stream()))
        cur = nxt
    return cur


# ------------------------------------------------------------------------------------ score net
class ScoreRunner:
    """ScoreNetwork lowered for a fixed (batch, length) and bound to device buffers."""

    def __init__(self, net, batch, t, device, shared=None):
        self.batch, self.t, self.device = batch, t, device
        with torch.no_grad():
            self.prog = P.lower_score_network(net, batch, t)
            self.proj = P.lower_cond_projection(net, batch, self.prog.meta["lengths"])
            self.exe = Executor(self.prog, device, shared=shared, tag="score")
            # the projection program writes straight into the score program's 'sc{lvl}' buffers
            self.proj_exe = Executor(self.proj, device,
                                     external=[n for n in self.proj.bufs if n.startswith("sc")],
                                     shared=shared, tag="proj")
            for name in self.proj.bufs:
                if name.startswith("sc"):
                    self.proj_exe.bufs[name] = self.exe.bufs[name]
            self._prepare_embedding(net, device)
        self.film = None
        self._film_tables = {}
        self.cond_lengths = self.prog.meta["lengths"]
        self.cond_channels = self.prog.meta["cond_channels"]

    def _prepare_embedding(self, net, device):
        self.dim = net.noise_cond_dim
        self.embed = embedding_spec(net.sigma_block, device)
        ws, bs = [], []
        for lin, off, cout in self.prog.film_layers:
            assert off == sum(w.shape[0] for w in ws)
            ws.append(fold.effective_weight(lin).float())
            bs.append(fold.bias_of(lin))
        self.film_w = torch.cat(ws).to(device).contiguous()
        self.film_b = torch.cat(bs).to(device).contiguous()
        self.film_cols = self.prog.film_cols

    def sigma_embedding(self, log10_sigma):
        """(rows,) fp32 log10 sigma -> (rows, noise_cond_dim) embedding g."""
        return sigma_embedding(self.embed, self.dim, log10_sigma)

    def set_sigmas(self, net_sigma):
        """net_sigma: (rows,) sigma values fed to the network (after the EDM noise scaling).
        Builds the FiLM table for all rows in two launches (embedding + one dense layer)."""
        g = self.sigma_embedding(torch.log10(net_sigma.float()))
        rows = g.shape[0]
        # persistent per row count: captured graphs keep pointing at it
        film = self._film_tables.get(rows)
        if film is None:
            film = torch.empty(rows, self.film_cols, dtype=torch.float32, device=self.device)
            self._film_tables[rows] = film
        lib.check(lib.load().ou_linear_f32(_ptr(g), _ptr(self.film_w), _ptr(self.film_b), _ptr(film),
                                           rows, self.dim, self.film_cols, self.film_cols, 0, 0.0,
                                           _stream()))
        self.film = film
        return film

    def set_cond(self, cond_blocked):
        """cond_blocked: list of blocked bf16 conditioning tensors (coarsest first).  Runs the
        step-invariant signal_cond_proj 1x1 convs once."""
        for lvl, c in enumerate(cond_blocked):
            cb = channel_block(self.cond_channels[lvl])
            want = (self.batch, self.cond_channels[lvl] // cb, self.cond_lengths[lvl], cb)
            if tuple(c.shape) != want:
                raise ValueError(f"conditioning tensor {lvl} has blocked shape {tuple(c.shape)}, "
                                 f"expected {want}")
            self.proj_exe.bufs[f"cond{lvl}"] = c
        self.proj_exe.run()

    def _build_plan(self):
        """Record the evaluation as an ``ou_plan`` (include/ou_b200.h, plan level): every op with its device
        pointers resolved, so that ``step`` is ONE foreign call instead of one per launch."""
        self._plan, self._plan_ctas = record_plan(self.exe), self.exe.max_ctas

    def __del__(self):
        plan = self.__dict__.get("_plan")
        if plan is not None:
            try:
                lib.load().ou_plan_destroy(plan)
            except Exception:
                pass

    def step(self, x, film_row, per_clip_film, in_scale=None, coef=None, noise=None, xout=None,
             net_out=None, ops=None, op_range=None):
        """One network evaluation (or the part ``ops`` / ``op_range`` = (first, count) of it).  x: (B,1,T)
        fp32.  film_row: first row of the FiLM table to use (one row shared by all clips, or B consecutive
        rows if per_clip_film)."""
        film = self.film[film_row:]
        bstride = self.film_cols if per_clip_film else 0
        if USE_PLAN and PROFILE is None and USE_TRUNK and not self.exe.naive and (ops is None or op_range is not None):
            if self.__dict__.get("_plan") is None or self._plan_ctas != self.exe.max_ctas:
                self._build_plan()
            args = lib.StepArgs()
            args.x, args.film, args.film_bstride = x.data_ptr(), film.data_ptr(), bstride
            args.in_scale = in_scale.data_ptr() if in_scale is not None else None
            args.coef = coef.data_ptr() if coef is not None else None
            args.noise = noise.data_ptr() if noise is not None else None
            args.xout = xout.data_ptr() if xout is not None else None
            args.net_out = net_out.data_ptr() if net_out is not None else None
            first, count = op_range if op_range is not None else (0, -1)
            lib.check(lib.load().ou_plan_run(self._plan, byref(args), first, count, _stream()))
            return
        self.exe.bufs["x"] = x
        self.exe.run(film=film, film_bstride=bstride,
                     in_scale=in_scale, coef=coef, noise=noise, xout=xout, net_out=net_out, ops=ops)

    def phases(self):
        """(ops up to and including the GRU input projection, the recurrence, the rest) -- the recurrence is
        a latency-bound kernel on a few SMs that ``PipelinedScoreRunner`` hides behind the other half-batch's
        convolutions; None when the network has no recurrent bottleneck."""
        ops = self.prog.ops
        idx = [i for i, op in enumerate(ops) if isinstance(op, P.GruOp)]
        if len(idx) != 1:
            return None
        i = idx[0]
        return ops[:i], ops[i:i + 1], ops[i + 1:]

    def phase_ranges(self):
        """(first, count) of the three phases in the plan's op numbering."""
        enc, gru, dec = self.phases()
        return (0, len(enc)), (len(enc), 1), (len(enc) + 1, len(dec))


# Two half-batch op streams per step (PipelinedScoreRunner) pay when the recurrence is a large share of a
# step, i.e. at small per-GPU batches: measured on B200 +8 % at B = 4 (cfg-4 share of a GPU), +2 % at
# B = 32 x 8 s (cfg-2), -4 % at B = 64 x 4 s (cfg-3: twice the launches for a recurrence half as long).
# OU_PIPELINE = auto (default: batch <= PIPELINE_MAX_BATCH) | 1 (always) | 0 (never, round-1 behaviour)
PIPELINE = os.environ.get("OU_PIPELINE", "auto")
PIPELINE_MAX_BATCH = 32


def _pipeline_wanted(batch):
    if PIPELINE == "0":
        return False
    return batch >= 2 and (PIPELINE == "1" or batch <= PIPELINE_MAX_BATCH)


class PipelinedScoreRunner:
    """The score network for a batch split in two halves that run as two op streams on two CUDA streams
    (SURVEY section 8(f) item 1).  Batch rows are independent end to end, so are the halves -- over the
    WHOLE sampler loop, not just one step.  The bottleneck BiGRU is a serial recurrence (801 steps x
    ~0.8 us) that keeps 8 SMs per (direction, 8 clips) busy whatever the batch: run alone it idles most
    of the GPU for ~15 % of every step.  Here half B's encoder is ordered after half A's (one event per
    step), which keeps the halves half a step apart: A's recurrence runs while B's encoder convolutions
    use the other SMs, B's while A's decoder does.  The persistent conv / trunk grids are capped at
    (SMs - CTAs of the other half's recurrence) so that both fit side by side (``max_ctas`` of the C
    ABI), and ``ou_gru_bidir`` launches at the highest priority."""

    def __init__(self, net, batch, t, device, shared=None):
        self.batch, self.t, self.device = batch, t, device
        b0 = (batch + 1) // 2
        self.rows = [(0, b0), (b0, batch)]
        self.halves = [ScoreRunner(net, hi - lo, t, device, shared) for lo, hi in self.rows]
        self.film_cols = self.halves[0].film_cols
        self.film = None
        n_sms = torch.cuda.get_device_properties(device).multi_processor_count
        # The overlapped recurrence is never on the critical path here: clusters of 4 CTAs (a step takes ~30 % longer,
        # half the SMs: measured 1 000 -> 1 025 audio-s/s at cfg-2) leave more SMs to the other half's convolutions
        hidden = [op.hidden for op in self.halves[0].prog.ops if isinstance(op, P.GruOp)]
        cluster = int(os.environ.get("OU_PIPE_GRU_CLUSTER", "4"))
        for h, other in ((0, 1), (1, 0)):
            self.halves[h].exe.gru_cluster = cluster
            gru_ctas = lib.load().ou_gru_ctas(hidden[0] if hidden else 256, self.halves[other].batch, cluster)
            self.halves[h].exe.max_ctas = max(n_sms - gru_ctas, n_sms // 2)
            if os.environ.get("OU_PIPE_CAP"):        # A/B knob: explicit cap (0 = none)
                self.halves[h].exe.max_ctas = int(os.environ["OU_PIPE_CAP"])
        self.side = torch.cuda.Stream(device=device)
        self.cond_lengths = self.halves[0].cond_lengths
        self.cond_channels = self.halves[0].cond_channels

    def set_sigmas(self, net_sigma):
        self.film = self.halves[0].set_sigmas(net_sigma)
        self.halves[1].film = self.film          # same table (one row per step), same device pointer
        return self.film

    def set_cond(self, cond_blocked):
        for h, (lo, hi) in zip(self.halves, self.rows):
            h.set_cond([c[lo:hi] for c in cond_blocked])

    def loop_step(self, n, x, in_scale, coef, noise):
        """Step ``n`` of the sampler for both halves; must be called with ``self.side`` already forked from
        the current stream (``SamplerLoop._loop``)."""
        main, side = torch.cuda.current_stream(), self.side
        a, b = self.halves
        (lo_a, hi_a), (lo_b, hi_b) = self.rows

        def part(h, lo, hi, ops, rng=None):
            h.step(x[lo:hi], n, False, in_scale=in_scale[lo:hi], coef=coef[lo:hi],
                   noise=None if noise is None else noise[lo:hi], xout=x[lo:hi], ops=ops, op_range=rng)

        enc_a, gru_a, dec_a = a.phases()
        r_enc, r_gru, r_dec = a.phase_ranges()
        part(a, lo_a, hi_a, enc_a, r_enc)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            if os.environ.get("OU_PIPE_STAGGER", "1") != "0":
                side.wait_event(ev)              # B's encoder after A's: the halves stay half a step apart
            part(b, lo_b, hi_b, None)
        part(a, lo_a, hi_a, gru_a, r_gru)
        part(a, lo_a, hi_a, dec_a, r_dec)


# OU_PLAN=0: one ctypes call per launch (Executor.run) instead of one ``ou_plan_run`` per evaluation
USE_PLAN = os.environ.get("OU_PLAN", "1") != "0"

# OU_GRAPH=0 launches the sampler loop kernel by kernel from Python instead of replaying a CUDA graph
USE_GRAPH = os.environ.get("OU_GRAPH", "1") != "0"
GRAPH_NOISE_LIMIT = 8 << 30     # bytes of pre-drawn noise above which the loop runs un-captured
GRAPH_KERNELS = 0               # kernels of this library executed through graph replays so far


def kernel_count():
    """Kernels of libou_b200.so executed so far: direct launches + kernel nodes of replayed graphs."""
    return lib.launch_count() + GRAPH_KERNELS


class SamplerLoop:
    """The N-step reverse-SDE loop of ``Universe.enhance`` (universe.py:334-343) for one
    ScoreRunner, captured ONCE as a CUDA graph (N x 34 kernel nodes) over persistent buffers:
    x, the per-step noise, FiLM table, input scales and update coefficients are static device
    tensors whose CONTENTS are refreshed per call, so replaying costs one launch instead of
    ~2 200 ctypes calls.  Noise is still drawn by ``torch.randn`` in the reference's call order
    (a5 of SURVEY section 8), just before the replay instead of inside the loop."""

    def __init__(self, sr, n_steps, n_start=0):
        self.sr, self.n_steps, self.n_start = sr, n_steps, n_start
        dev, B, T = sr.device, sr.batch, sr.t
        self.x = torch.empty(B, 1, T, dtype=torch.float32, device=dev)
        self.noise = torch.empty(max(n_steps - 1, 1), B, 1, T, dtype=torch.float32, device=dev)
        self.in_scale = torch.ones(n_steps, B, dtype=torch.float32, device=dev)
        self.coef = torch.zeros(n_steps, B, 3, dtype=torch.float32, device=dev)
        self.graph = None

    def _loop(self):
        sr, N = self.sr, self.n_steps
        if isinstance(sr, PipelinedScoreRunner):
            main = torch.cuda.current_stream()
            sr.side.wait_stream(main)            # fork (inside a capture: the side stream joins it)
            for n in range(self.n_start, N):
                sr.loop_step(n, self.x, self.in_scale[n], self.coef[n],
                             self.noise[n] if n < N - 1 else None)
            main.wait_stream(sr.side)            # join
            return
        for n in range(self.n_start, N):   # warm start skips the first steps (universe.py:326-334)
            sr.step(self.x, n, False, in_scale=self.in_scale[n], coef=self.coef[n],
                    noise=self.noise[n] if n < N - 1 else None, xout=self.x)

    def capture(self):
        """Eager warm-up (lazy kernel attributes, tensor-map entry points) on a side stream, then
        capture.  Runs on whatever is in the static buffers: call before loading real inputs."""
        self.x.zero_()
        self.noise.zero_()
        side = torch.cuda.Stream(device=self.sr.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            if isinstance(self.sr, PipelinedScoreRunner):
                self.sr.side.wait_stream(side)
                self.sr.loop_step(0, self.x, self.in_scale[0], self.coef[0], None)
                side.wait_stream(self.sr.side)
            else:
                self.sr.step(self.x, 0, False, in_scale=self.in_scale[0], coef=self.coef[0],
                             noise=None, xout=self.x)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        n0 = lib.launch_count()
        with torch.cuda.graph(g):
            self._loop()
        self.n_kernels = lib.launch_count() - n0      # kernel nodes of the graph
        self.graph = g

    def run(self):
        global GRAPH_KERNELS
        if self.graph is not None:
            self.graph.replay()
            GRAPH_KERNELS += self.n_kernels
        else:
            self._loop()


def get_sampler_loop(sr, n_steps, n_start=0):
    """Cached SamplerLoop of a ScoreRunner; captured as a CUDA graph unless disabled, profiled
    (bench.py's per-launch events) or too large."""
    loops = sr.__dict__.setdefault("_loops", OrderedDict())
    want_graph = (USE_GRAPH and PROFILE is None and
                  (n_steps - 1) * sr.batch * sr.t * 4 <= GRAPH_NOISE_LIMIT)
    key = (n_steps, n_start, want_graph)
    if key in loops:
        loops.move_to_end(key)
        return loops[key]
    while len(loops) >= MAX_CACHED_LOOPS:     # each loop owns an (N-1) x B x T fp32 noise buffer + a graph
        loops.popitem(last=False)
    loop = SamplerLoop(sr, n_steps, n_start)
    if want_graph:
        if sr.film is None or sr.film.shape[0] < n_steps:
            raise RuntimeError("set_sigmas() must run before the sampler loop is captured")
        loop.capture()
    loops[key] = loop
    return loop


def get_score_runner(net, batch, t, device, pipelined=False):
    """``pipelined``: the sampler loop's runner -- two half-batch op streams (``PipelinedScoreRunner``) when
    the batch can be split and the network has a recurrent bottleneck; single evaluations
    (``score_forward``, the EDM wrapper) and bench.py's kernel-by-kernel timing pass use the plain one."""
    device = torch.device(device)
    profiling = PROFILE is not None and not os.environ.get("OU_PIPE_PROFILE")   # tools/pipe_timeline.py
    if pipelined and _pipeline_wanted(batch) and not profiling and getattr(net.encoder, "seq_model", "") == "gru":
        return _get_runner(net, ("score2", batch, t, str(device)),
                           lambda shared: PipelinedScoreRunner(net, batch, t, device, shared))
    return _get_runner(net, ("score", batch, t, str(device)),
                       lambda shared: ScoreRunner(net, batch, t, device, shared))


@on_tensor_device
def score_forward(net, x, sigma, cond):
    """ScoreNetwork.forward(x, sigma, cond) on reference-layout tensors (score.py:277-297)."""
    require_cuda(x, sigma, *cond)
    same_device(x, sigma, *cond)
    b, c, t = x.shape
    if c != 1:
        raise ValueError("ScoreNetwork expects a (B, 1, T) input")
    r = get_score_runner(net, b, t, x.device)
    with torch.no_grad():
        r.set_sigmas(sigma.reshape(-1))
        r.set_cond([pack_blocked(ci) for ci in cond])
        out = torch.empty(b, 1, t, dtype=torch.float32, device=x.device)
        r.step(x.contiguous().float(), 0, True, net_out=out)
    return out


# ------------------------------------------------------------------------------------ conditioner
class ConditionerRunner:
    def __init__(self, net, batch, t, device, need_signal_tail=True, shared=None):
        self.batch, self.t, self.device = batch, t, device
        with torch.no_grad():
            self.prog = P.lower_conditioner(net, batch, t, need_signal_tail)
            self.exe = Executor(self.prog, device, shared=shared, tag="cond")
        self.n_cond = len([k for k in self.prog.outputs if k.startswith("cond")])

    def run(self, x, x_wav=None):
        """ConditionerNetwork.forward as ONE foreign call (``ou_plan_run`` over the recorded plan: SURVEY 8b's
        ``ou_condition_forward``); op by op through the Executor when profiling or with OU_PLAN=0."""
        if USE_PLAN and PROFILE is None and USE_TRUNK and not self.exe.naive:
            if self.__dict__.get("_plan") is None:
                self._plan = record_plan(self.exe)
            args = lib.StepArgs()
            args.x = x.data_ptr()
            args.x_wav = None if x_wav is None else x_wav.data_ptr()
            lib.check(lib.load().ou_plan_run(self._plan, byref(args), 0, -1, _stream()))
        else:
            self.exe.bufs["x"] = x
            self.exe.bufs["x_wav"] = x if x_wav is None else x_wav
            self.exe.run()
        out = self.prog.outputs
        cond = [self.exe.bufs[out[f"cond{i}"]] for i in range(self.n_cond)]
        y_hat = self.exe.bufs[out["y_hat"]] if "y_hat" in out else None
        return cond, y_hat, self.exe.bufs[out["h"]]


    def __del__(self):
        plan = self.__dict__.get("_plan")
        if plan is not None:
            try:
                lib.load().ou_plan_destroy(plan)
            except Exception:
                pass


def get_conditioner_runner(net, batch, t, device, need_signal_tail=True):
    device = torch.device(device)
    return _get_runner(net, ("cond", batch, t, str(device), need_signal_tail),
                       lambda shared: ConditionerRunner(net, batch, t, device, need_signal_tail, shared))


@on_tensor_device
def conditioner_forward(net, x, x_wav=None):
    """ConditionerNetwork.forward on reference-layout tensors -> (conditions, y_hat, h) fp32."""
    require_cuda(x, x_wav)
    b, c, t = x.shape
    if c != 1:
        raise ValueError("ConditionerNetwork expects a (B, 1, T) input")
    r = get_conditioner_runner(net, b, t, x.device, True)
    with torch.no_grad():
        cond, y_hat, h = r.run(x.contiguous().float(),
                               None if x_wav is None else x_wav.contiguous().float())
        cond = [unpack_blocked(ci) for ci in cond]
        y_hat = unpack_blocked(y_hat)
        y_hat = torch.nn.functional.pad(y_hat, (0, t - y_hat.shape[-1]))
        return cond, y_hat, unpack_blocked(h)


@on_tensor_device
def compute_mel_spec(mel_adapter, x):
    """MelAdapter.compute_mel_spec (condition.py:92-108): (B,1,T) -> (B, n_mels, frames) fp32."""
    require_cuda(x)
    b, _, t = x.shape
    hop = mel_adapter.ds_factor
    frames = P.ceil_div(t, hop)
    prog = P.Program(b)
    op = P.MelOp("mel", "x", "mel", mel_adapter.n_fft, hop, mel_adapter.n_mels, mel_adapter.pad_left,
                 frames, t, mel_adapter.mel_spec.spectrogram.window.detach().float(),
                 mel_adapter.mel_spec.mel_scale.fb.detach().float())
    prog.ops.append(op)
    prepare_ops(prog, x.device)
    pk = op.packed
    L = lib.load()
    xc = x.contiguous().float()
    lib.check(L.ou_mel_power(_ptr(xc), _ptr(pk["window"]), _ptr(pk["fb"]), _ptr(pk["dft"]),
                             _ptr(pk["power"]), _ptr(pk["mel"]), _ptr(pk["energy"]), b, t, op.n_fft, hop,
                             op.n_mels,
                             op.pad_left, frames, _stream()))
    lib.check(L.ou_mel_finalize(_ptr(pk["mel"]), _ptr(pk["energy"]), _ptr(pk["mel"]), None, b,
                                op.n_mels, frames, _stream()))
    return pk["mel"]


# ------------------------------------------------------------------------------------ block-level
@on_tensor_device
def conv_block_forward(blk, h, noise_cond=None, input_cond=None, res=None, length=None):
    """ConvBlock.forward on (B, C, T) fp32 tensors -> (h_out, skip, cond_out)  (blocks.py:327-412).
    Layer-level entry point (parity tests, LoRA-style consumers); the networks never call it."""
    require_cuda(h, noise_cond, input_cond, res)
    b, c, t = h.shape
    prog = P.Program(b)
    prog.buf("in", "blocked", c, t)
    src = "in"
    bufs_in = {"in": pack_blocked(h)}
    if res is not None and blk.rate_change_dir == "none":
        # rate-preserving block: the network lowering folds this add into the producer; do it here
        bufs_in["in"] = pack_blocked((h + res) * P.SQRT_HALF)
        res = None
    film_lin = None
    if noise_cond is not None:
        film_lin = _FilmPassthrough(blk.n_channels)
    if input_cond is not None:
        prog.buf("icond", "blocked", blk.n_channels, input_cond.shape[-1])
        bufs_in["icond"] = pack_blocked(input_cond)
    if res is not None:
        prog.buf("res", "blocked", blk.n_channels, res.shape[-1])
        bufs_in["res"] = pack_blocked(res)
    with torch.no_grad():
        out, _, skip, cond = P.lower_conv_block(
            prog, blk, "blk", src, t, film_linear=film_lin,
            input_cond="icond" if input_cond is not None else None,
            res="res" if res is not None else None, length=length, raw_cond_out=True)
        exe = Executor(prog, h.device, external=list(bufs_in))
        exe.bufs.update(bufs_in)
        film = noise_cond.contiguous().float() if noise_cond is not None else None
        exe.run(film=film, film_bstride=2 * blk.n_channels)
        return (unpack_blocked(exe.bufs[out]), unpack_blocked(exe.bufs[skip]),
                unpack_blocked(exe.bufs[cond]))


class _FilmPassthrough:
    """Marker so that ``lower_conv_block`` allocates FiLM columns when the caller passes the
    already-projected (B, 2C) vector (ConvBlock.forward's ``noise_cond``)."""

    def __init__(self, c):
        self.c = c


@on_tensor_device
def alias_free_snake(mod, x, conv=None, blocked=False, t=None):
    """AliasFreeSnake.forward (bigvgan/snake.py:127-157), optionally fused with the k-tap conv to one
    channel behind it (``conv``: a Conv1d with out_channels == 1, 'same' padding).
    x: (B, C, T) fp32, or a blocked bf16 activation buffer when ``blocked``.
    Returns (B, C, T) fp32 without ``conv``, (B, 1, T) fp32 with it."""
    require_cuda(x)
    act = mod.act
    if act.up_ratio != 2 or act.down_ratio != 2:
        raise NotImplementedError("AliasFreeSnake kernel covers the x2 / :2 configuration upstream uses")
    snake = act.act
    if blocked:
        b, nblk, tt, cb = x.shape
        c = nblk * cb
    else:
        b, c, tt = x.shape
        x = x.contiguous().float()
    if t is not None and t != tt:
        raise ValueError(f"expected length {t}, got {tt}")
    dev = x.device
    alpha = snake.alpha.detach().float().to(dev).contiguous()
    beta = snake.beta.detach().float().to(dev).contiguous() if hasattr(snake, "beta") else None
    ku = act.upsample.kernel.detach().float().to(dev).contiguous()      # (2, 1, up_len)
    kd = act.downsample.kernel.detach().float().to(dev).contiguous()    # (1, 1, down_len)
    if ku.shape[0] != 2 or kd.shape[0] != 1:
        raise ValueError("unexpected resampling kernel shapes")
    w = bias = None
    k = 1
    if conv is not None:
        if fold.inner(conv).out_channels != 1 or fold.inner(conv).in_channels != c:
            raise ValueError("the fused conv must map all channels to one")
        w = fold.effective_weight(conv)[0].float().to(dev).contiguous()   # (C, k)
        k = w.shape[1]
        bias = fold.bias_of(conv)
        bias = float(bias[0].item()) if bias is not None else 0.0
        out = torch.empty(b, 1, tt, dtype=torch.float32, device=dev)
    else:
        out = torch.empty(b, c, tt, dtype=torch.float32, device=dev)
    lib.check(lib.load().ou_alias_free_snake(
        _ptr(x), 1 if blocked else 0, _ptr(alpha), _ptr(beta), 1 if snake.alpha_logscale else 0,
        _ptr(ku), ku.shape[-1], _ptr(kd), kd.shape[-1], _ptr(w), bias or 0.0, k, _ptr(out), b, c, tt,
        _stream()))
    return out


@on_tensor_device
def prelu_conv_forward(pc, x, blocked=False):
    """PReLU_Conv.forward on a (B, C, T) fp32 tensor (blocks.py:205-227)."""
    require_cuda(x)
    if pc.act_type in ("snake", "snakebeta"):
        # only instance upstream: UniverseGAN.signal_decoupling_layer (universe_gan.py:117-126)
        if (pc.out_channels != 1 or pc.stride != 1 or pc.padding != "same" or pc.use_transpose
                or pc.antialiasing):
            raise NotImplementedError("snake-activated PReLU_Conv is only built as the signal "
                                      "decoupling layer (C -> 1, 'same')")
        return alias_free_snake(pc.prelu, x, conv=pc.conv, blocked=blocked)
    if blocked:
        raise NotImplementedError("blocked input is only accepted by the signal decoupling layer")
    if pc.act_type != "prelu":
        raise NotImplementedError("only act_type='prelu' has a kernel")
    b, c, t = x.shape
    if pc.stride == 1 and pc.padding != "same":
        raise NotImplementedError("stride-1 PReLU_Conv is only used with padding='same' upstream")
    prog = P.Program(b)
    prog.buf("in", "blocked", c, t)
    with torch.no_grad():
        fc = fold.fold_prelu_conv(pc)
        dst, _ = P.add_conv(prog, "pc", "in", "out", fc, t)
        exe = Executor(prog, x.device, external=["in"])
        exe.bufs["in"] = pack_blocked(x)
        exe.run()
        return unpack_blocked(exe.bufs[dst])


@on_tensor_device
def film(x, y):
    """film(x, y) = y[:, :C] * x + y[:, C:]  (blocks.py:53-59) on (B, C, T) / (B, 2C) fp32 tensors."""
    require_cuda(x, y)
    if y.shape[1] != 2 * x.shape[1]:
        raise ValueError("g should have 2 times more channels than y")
    b, c, t = x.shape
    out = torch.empty_like(x, dtype=torch.float32)
    xc, yc = x.contiguous().float(), y.contiguous().float()   # keep alive across the launch
    lib.check(lib.load().ou_film_f32(_ptr(xc), _ptr(yc), _ptr(out), b, c, t, _stream()))
    return out


def lowpass(x, taps):
    raise NotImplementedError(
        "BinomialAntiAlias is folded into the rate-change conv weights (engine.fold); it has no "
        "standalone kernel")


@on_tensor_device
def sigma_embed(block, log10_sigma):
    """sigma_block(log10 sigma) -> (B, noise_cond_dim) fp32 (sigma_block.py:50-57, 73-78)."""
    require_cuda(log10_sigma)
    with torch.no_grad():
        return sigma_embedding(embedding_spec(block, log10_sigma.device), block.n_dim,
                               log10_sigma.reshape(-1))
