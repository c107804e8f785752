"""Per-kernel parity on the GPU: every C-ABI entry point against the PyTorch emulator / oracle.

Tolerances: bit-level agreement is not expected for floating point -- the tensor-core kernel
accumulates in a different order than ATen.  The kernel and the emulator round to the storage
type (fp16 by default) at the SAME points, so they agree to ~1e-3 of the tensor RMS (a few ulp flips); fp32-only kernels
agree to ~1e-5.
"""
import math

import pytest
import torch

from common import rel_rms
from oracle import emulator as E
from oracle import universe_oracle as O
from open_universe_b200.engine import lib, program as P, runtime as R
from open_universe_b200.engine.fold import FoldedConv

pytestmark = pytest.mark.gpu
DEV = "cuda"


E.QDTYPE = lib.act_dtype()     # the emulator rounds where and how the loaded library does


def bf(x):
    return x.to(E.QDTYPE).float()


def rand_fc(g, cin, cout, s=1, up=1, taps=1, tap_off=0, prelu_in=None):
    w = torch.randn(up * cout, taps, s * cin, generator=g) / math.sqrt(taps * s * cin)
    return FoldedConv(w, 0.1 * torch.randn(up * cout, generator=g), cin, cout, s, up, taps, tap_off,
                      prelu_in)


CONV_CASES = [
    # cin, cout, s, up, taps, off, t_in, B, options
    dict(cin=32, cout=32, taps=5, off=-2, t=1000, B=2, prelu_in=0.2, film=True, prelu_out=0.3),
    dict(cin=64, cout=64, taps=3, off=-1, t=777, B=1, add1=True, prelu_out=0.1, prelu_out2=-0.2),
    dict(cin=128, cout=128, taps=5, off=-2, t=300, B=3, add1=True, add2=True),
    dict(cin=512, cout=512, taps=3, off=-1, t=51, B=2, prelu_in=0.25),
    dict(cin=32, cout=64, s=2, taps=3, off=-1, t=1001, B=2, prelu_in=0.25),          # anti-aliased down
    dict(cin=256, cout=512, s=5, taps=1, off=0, t=403, B=1, prelu_in=0.3),           # plain down
    dict(cin=64, cout=32, up=2, taps=3, off=-1, t=500, B=2, prelu_in=0.2, add1=True, t_out=999),
    dict(cin=512, cout=256, up=5, taps=1, off=0, t=40, B=1, add1=True),              # plain up
    dict(cin=32, cout=512, s=160, taps=1, off=0, t=3200, B=2, prelu_in=0.25, add1=True),  # st_conv
    dict(cin=80, cout=512, taps=3, off=-1, t=60, B=1),                               # mel conv (K pad)
    dict(cin=48, cout=96, s=2, taps=3, off=-1, t=601, B=1, prelu_in=0.25),           # 24 kHz widths
    dict(cin=96, cout=48, up=2, taps=3, off=-1, t=300, B=1, add1=True),
    dict(cin=512, cout=1536, taps=1, off=0, t=100, B=2, f32_tm=True),                # GRU x-proj
    # multi-tile persistence / weight residency / weight streaming of the tcgen05 kernel
    dict(cin=32, cout=32, taps=5, off=-2, t=128160, B=2, prelu_in=0.25, add1=True),
    dict(cin=64, cout=64, taps=3, off=-1, t=64080, B=1, prelu_out=0.25),
    dict(cin=128, cout=128, taps=5, off=-2, t=16020, B=2, prelu_in=0.25, film=True, add1=True),
    dict(cin=256, cout=256, taps=5, off=-2, t=4005, B=8, prelu_in=0.25),
    dict(cin=512, cout=512, taps=3, off=-1, t=801, B=16, add1=True),
    dict(cin=512, cout=256, up=5, taps=3, off=-1, t=801, B=3, prelu_in=0.2, add1=True),
    dict(cin=512, cout=1536, taps=1, off=0, t=801, B=4, f32_tm=True),
    # 2 / 4 sub-tiles of 128 rows per scheduling unit (the planner needs >= 4 units per SM):
    # bn = 64 strided down conv, bn = 64 up conv with skip, bn = 128 streamed and resident weights
    dict(cin=32, cout=64, s=2, taps=3, off=-1, t=128160, B=5, prelu_in=0.25),
    dict(cin=64, cout=32, up=2, taps=3, off=-1, t=64080, B=5, prelu_in=0.25, add1=True),
    dict(cin=64, cout=128, s=4, taps=3, off=-1, t=64080, B=10, prelu_in=0.25),
    dict(cin=128, cout=128, taps=3, off=-1, t=16020, B=10, add1=True),
    dict(cin=128, cout=128, taps=5, off=-2, t=16020, B=10, prelu_in=0.25, film=True),
]


@pytest.mark.parametrize("c", CONV_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv1d_vs_emulator(c):
    g = torch.Generator().manual_seed(1234)
    B, cin, cout, t = c["B"], c["cin"], c["cout"], c["t"]
    s, up = c.get("s", 1), c.get("up", 1)
    fc = rand_fc(g, cin, cout, s, up, c["taps"], c["off"], c.get("prelu_in"))
    prog = P.Program(B)
    prog.buf("in", "blocked", cin, t)
    kw = {}
    t_out = c.get("t_out")
    if c.get("f32_tm"):
        kw["dst_kind"] = "f32_blk"
    dst, t_out = P.add_conv(prog, "c", "in", "out", fc, t, t_out, **kw)
    op = prog.ops[0]
    x = bf(torch.randn(B, cin, t, generator=g))
    inputs = {"in": x}
    for name in ("add1", "add2"):
        if c.get(name):
            prog.buf(name, "blocked", cout, t_out)
            setattr(op, name, name)
            inputs[name] = bf(torch.randn(B, cout, t_out, generator=g))
    op.scale1, op.scale2 = 0.7071, 0.5 if c.get("add2") else 1.0
    film = None
    if c.get("film"):
        op.film_off = 0
        film = torch.randn(B, 2 * cout, generator=g)
    op.prelu_out, op.prelu_out2 = c.get("prelu_out"), c.get("prelu_out2")

    bufs, _, _ = E.run_program(prog, inputs, film=film, quant=True)
    want = bufs["out"]

    exe = R.Executor(prog, DEV, external=list(inputs))
    for k, v in inputs.items():
        exe.bufs[k] = R.pack_blocked(v.to(DEV))
    film_d = film.to(DEV).contiguous() if film is not None else None
    for naive in (True, False):
        exe.naive = naive
        exe.run(film=film_d, film_bstride=2 * cout)
        got = exe.bufs["out"]
        got = R.unpack_f32_blocked(got).cpu() if c.get("f32_tm") else R.unpack_blocked(got).cpu()
        assert got.shape == want.shape
        err = rel_rms(got, want)
        if err >= 3e-3:   # locate the damage: which clip / channel / time the worst element sits at
            d = (got - want).abs()
            idx = [int(i) for i in torch.unravel_index(d.argmax(), d.shape)]
            bad = (d > 0.05 * want.abs().max()).float()
            raise AssertionError(f"naive={naive} rel_rms={err:.4f} worst at {idx} got={got[tuple(idx)]:.4f} "
                                 f"want={want[tuple(idx)]:.4f} bad_fraction={bad.mean():.4f} "
                                 f"bad_per_batch={bad.flatten(1).mean(1).tolist()}")


TRUNK_CASES = [
    dict(C=64, t=1000, B=2, sc=True, film=True, po=(0.3, None)),
    dict(C=32, t=1000, B=2, sc=True, film=True, po=(0.3, -0.2)),
    dict(C=64, t=123, B=1, sc=False, film=False, po=(None, None)),      # single partial item
    dict(C=32, t=251, B=3, sc=False, film=True, po=(None, None)),       # one short of an item
    dict(C=64, t=124 * 7, B=1, sc=True, film=False, po=(None, None)),   # exact multiple of the item
    dict(C=32, t=252 * 3 + 1, B=2, sc=True, film=False, po=(0.1, None)),
    dict(C=64, t=64080, B=3, sc=True, film=True, po=(None, None)),      # many items per CTA, 3 slots
    dict(C=32, t=128160, B=2, sc=False, film=True, po=(0.25, 0.25)),
]


@pytest.mark.parametrize("c", TRUNK_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_trunk_vs_emulator(c):
    """ou_conv_trunk against the emulator AND against its own three-launch expansion."""
    g = torch.Generator().manual_seed(4321)
    B, C, t = c["B"], c["C"], c["t"]
    prog = P.Program(B)
    prog.buf("in", "blocked", C, t)
    inputs = {"in": bf(torch.randn(B, C, t, generator=g))}
    if c["sc"]:
        prog.buf("sc", "blocked", C, t)
        inputs["sc"] = bf(torch.randn(B, C, t, generator=g))
    fc1 = rand_fc(g, C, C, taps=5, tap_off=-2, prelu_in=0.2)
    fc2 = rand_fc(g, C, C, taps=3, tap_off=-1)
    fc3 = rand_fc(g, C, C, taps=3, tap_off=-1)
    P.add_conv(prog, "conv1", "in", "c1", fc1, t, add1="sc" if c["sc"] else None,
               scale1=0.7071 if c["sc"] else 1.0, film_off=0 if c["film"] else None, prelu_out=0.15)
    P.add_conv(prog, "conv2", "c1", "c2", fc2, t, prelu_out=-0.3)
    P.add_conv(prog, "conv3", "c2", "v", fc3, t, add1="in", scale1=0.7071, prelu_out=c["po"][0],
               prelu_out2=c["po"][1])
    assert P.fuse_trunk(prog, "trunk") and len(prog.ops) == 1
    film = torch.randn(B, 2 * C, generator=g) if c["film"] else None
    bufs, _, _ = E.run_program(prog, inputs, film=film, quant=True)
    want = bufs["v"]

    exe = R.Executor(prog, DEV, external=list(inputs))
    for k, v in inputs.items():
        exe.bufs[k] = R.pack_blocked(v.to(DEV))
    film_d = film.to(DEV).contiguous() if film is not None else None
    n0 = lib.launch_count()
    exe.run(film=film_d, film_bstride=2 * C)
    assert lib.launch_count() - n0 == 1          # the fused kernel ran, not the expansion
    got = R.unpack_blocked(exe.bufs["v"]).cpu()
    R.USE_TRUNK = False
    try:
        exe.bufs["v"].zero_()
        n0 = lib.launch_count()
        exe.run(film=film_d, film_bstride=2 * C)
        assert lib.launch_count() - n0 == 3
        split = R.unpack_blocked(exe.bufs["v"]).cpu()
    finally:
        R.USE_TRUNK = True
    assert torch.isfinite(got).all()
    err, err_split = rel_rms(got, want), rel_rms(got, split)
    if err >= 3e-3 or err_split >= 3e-3:
        d = (got - want).abs()
        idx = [int(i) for i in torch.unravel_index(d.argmax(), d.shape)]
        bad = (d > 0.05 * want.abs().max()).float()
        raise AssertionError(f"rel_rms={err:.4f} vs split {err_split:.4f} worst at {idx} got={got[tuple(idx)]:.4f} "
                             f"want={want[tuple(idx)]:.4f} bad_fraction={bad.mean():.5f} "
                             f"bad rows (time) {bad.amax(dim=(0, 1)).nonzero().flatten()[:20].tolist()}")


@pytest.mark.parametrize("hidden,B,T,add", [(256, 5, 37, True), (128, 2, 20, False),
                                            (384, 3, 25, True), (256, 32, 801, True),
                                            # batch <= 4: 4 clip slots per cluster; H = 384: cluster of 12
                                            (384, 4, 61, True), (384, 9, 33, False), (256, 3, 40, False),
                                            (128, 4, 9, True), (256, 1, 1, True), (384, 2, 3, False)])
def test_gru_vs_explicit(hidden, B, T, add):
    g = torch.Generator().manual_seed(7)
    H = hidden
    gx = torch.randn(B, T, 6 * H, generator=g)
    w_hh = torch.randn(2, 3 * H, H, generator=g) / math.sqrt(H)
    b_hh = 0.1 * torch.randn(2, 3 * H, generator=g)
    addt = bf(torch.randn(B, 2 * H, T, generator=g)) if add else None
    op = P.GruOp("g", "gx", "out", w_hh, b_hh, H, T, add="add" if add else None, scale=0.7071)
    bufs = {"gx": gx}
    if add:
        bufs["add"] = addt
    E.run_gru(op, bufs, quant=False)
    want = bufs["out"]
    out = R.alloc_blocked(B, 2 * H, T, DEV)
    # keep every device tensor referenced until the kernel has run (raw pointers are passed)
    d_gx, d_w, d_b = R.pack_f32_blocked(gx.to(DEV)), w_hh.to(DEV), b_hh.to(DEV)
    d_add = R.pack_blocked(addt.to(DEV)) if add else None
    lib.check(lib.load().ou_gru_bidir(R._ptr(d_gx), R._ptr(d_w), R._ptr(d_b), R._ptr(d_add), 0.7071,
                                      R._ptr(out), B, T, H, R._stream()))
    torch.cuda.synchronize()
    got = R.unpack_blocked(out).cpu()
    assert rel_rms(got, want) < 3e-3   # bf16 output rounding only
    # clusters of 4 CTAs (the SM-saving form the pipelined sampler asks for): same arithmetic, same bits
    out4 = R.alloc_blocked(B, 2 * H, T, DEV)
    assert lib.load().ou_gru_ctas(H, B, 4) == 2 * -(-B // (4 if B <= 4 else 8)) * (4 if H <= 256 else 8)
    lib.check(lib.load().ou_gru_bidir_ex(R._ptr(d_gx), R._ptr(d_w), R._ptr(d_b), R._ptr(d_add), 0.7071,
                                         R._ptr(out4), B, T, H, 4, R._stream()))
    torch.cuda.synchronize()
    assert torch.equal(out4, out)


@pytest.mark.parametrize("fs_cfg", [dict(n_fft=640, hop=160, n_mels=80), dict(n_fft=960, hop=240, n_mels=128)])
@pytest.mark.parametrize("T", [3200, 5003])
def test_mel_vs_oracle(fs_cfg, T):
    g = torch.Generator().manual_seed(3)
    B = 2
    x = 0.05 * torch.randn(B, 1, T, generator=g)
    rates = [2, 4, 4, 5] if fs_cfg["hop"] == 160 else [2, 3, 5, 8]
    cfg = dict(rate_factors=rates, n_mel_oversample=4, n_mels=fs_cfg["n_mels"])
    want = O.mel_spec(cfg, x)

    class Mel:  # minimal MelAdapter view
        pass
    from open_universe_b200.networks.universe.condition import MelAdapter
    m = MelAdapter(fs_cfg["n_mels"], 64, fs_cfg["hop"], 4).to(DEV)
    got = m.compute_mel_spec(x.to(DEV)).cpu()
    assert got.shape == want.shape
    assert rel_rms(got, want) < 2e-5


@pytest.mark.parametrize("T,tot", [(8000, 160), (4800, 160), (7000, 240)])
def test_pad_normalize_and_unpad(T, tot):
    g = torch.Generator().manual_seed(5)
    B = 3
    mix = 0.3 * torch.randn(B, 1, T, generator=g) + 0.01
    pad = tot - T % tot
    t_pad = T + pad
    level = 10 ** (-26 / 20)
    xp = torch.nn.functional.pad(mix, (pad // 2, pad - pad // 2))
    xp = xp - xp.mean(dim=(1, 2), keepdim=True)
    want = xp * (level / xp.std(dim=(1, 2), keepdim=True).clamp(min=1e-5))
    out = torch.empty(B, 1, t_pad, device=DEV)
    L = lib.load()
    d_mix = mix.to(DEV)
    lib.check(L.ou_pad_normalize(R._ptr(d_mix), R._ptr(out), None, B, T, t_pad, pad // 2, level,
                                 R._stream()))
    torch.cuda.synchronize()
    assert rel_rms(out.cpu(), want) < 1e-5
    # unpad + keep_rms + limiter
    x = 30.0 * want * torch.tensor([1.0, 0.001, 100.0])[:, None, None]
    mix_rms = mix.square().mean(dim=(-2, -1)).sqrt()
    for keep in (False, True):
        y = x[..., pad // 2: -(pad - pad // 2)]
        if keep:
            y = y * (mix_rms[:, None, None] / y.square().mean(dim=(-2, -1), keepdim=True).sqrt().clamp(min=1e-5))
        sc = y.abs().max(dim=-1, keepdim=True).values
        y = torch.where(sc > 1.0, y / sc, y)
        o = torch.empty(B, 1, T, device=DEV)
        d_x, d_rms = x.to(DEV).contiguous(), mix_rms.to(DEV)
        lib.check(L.ou_unpad_limit(R._ptr(d_x), R._ptr(d_rms) if keep else None, R._ptr(o), B, t_pad,
                                   pad // 2, T, T, R._stream()))
        torch.cuda.synchronize()
        assert rel_rms(o.cpu(), y) < 1e-5


def test_input_and_output_kernels():
    g = torch.Generator().manual_seed(9)
    B, T, C = 2, 1003, 32
    x = torch.randn(B, 1, T, generator=g)
    w = torch.randn(C, 3, generator=g)
    b = torch.randn(C, generator=g)
    sc = torch.tensor([0.5, 2.0])
    want = torch.nn.functional.conv1d(x * sc[:, None, None], w[:, None, :], b, padding="same")
    out = R.alloc_blocked(B, C, T, DEV)
    L = lib.load()
    d_x, d_w, d_b, d_sc = x.to(DEV), w.to(DEV), b.to(DEV), sc.to(DEV)
    lib.check(L.ou_input_conv(R._ptr(d_x), R._ptr(d_w), R._ptr(d_b), R._ptr(d_sc), R._ptr(out), B, T, C,
                              3, R._stream()))
    torch.cuda.synchronize()
    assert rel_rms(R.unpack_blocked(out).cpu(), bf(want)) < 1e-3
    # output conv + SDE update
    src = bf(torch.randn(B, C, T - 3, generator=g))
    wo = torch.randn(C, 3, generator=g) / 10
    coef = torch.randn(B, 3, generator=g)
    noise = torch.randn(B, 1, T, generator=g)
    net = torch.nn.functional.conv1d(src, wo[None], None, padding="same") + 0.3
    net = torch.nn.functional.pad(net, (0, 3))
    xnew = coef[:, 0, None, None] * x + coef[:, 1, None, None] * net + coef[:, 2, None, None] * noise
    xo = torch.empty(B, 1, T, device=DEV)
    no = torch.empty(B, 1, T, device=DEV)
    d_src, d_wo, d_coef, d_noise = R.pack_blocked(src.to(DEV)), wo.to(DEV), coef.to(DEV), noise.to(DEV)
    lib.check(L.ou_output_sde(R._ptr(d_src), R._ptr(d_wo), 0.3, R._ptr(d_coef), R._ptr(d_x), R._ptr(d_noise),
                              R._ptr(xo), R._ptr(no), B, C, 3, T - 3, T, R._stream()))
    torch.cuda.synchronize()
    assert rel_rms(no.cpu(), net) < 1e-5
    assert rel_rms(xo.cpu(), xnew) < 1e-5


def test_sigma_embeddings_and_linear():
    from open_universe_b200.networks.universe.sigma_block import SigmaBlock, SimpleTimeEmbedding
    torch.manual_seed(0)
    ls = torch.log10(torch.tensor([5.0, 0.3, 0.0005, 0.0125]))
    s = SimpleTimeEmbedding(512)
    with torch.no_grad():
        s.weight.fill_(0.7)
        s.bias.fill_(0.2)
    sd = {"p.weight": s.weight.detach(), "p.bias": s.bias.detach()}
    want = O.sigma_embedding({"time_embedding": "simple", "noise_cond_dim": 512}, sd, "p", ls)
    got = s.to(DEV)(ls.to(DEV)).cpu()
    assert (got - want).abs().max() < 2e-4      # fp32 phase up to ~800 rad: 1 ulp of phase = 6e-5
    r = SigmaBlock(32, 512)
    sd = {"p." + k: v.detach() for k, v in r.state_dict().items()}
    want = O.sigma_embedding({"noise_cond_dim": 512}, sd, "p", ls)
    got = r.to(DEV)(ls.to(DEV)).cpu()
    assert rel_rms(got, want) < 1e-3


def test_cpu_tensors_raise():
    from open_universe_b200.config import builtin_config, instantiate
    with pytest.raises(R.NoCudaPathError):
        R.pack_blocked(torch.zeros(1, 8, 4))


@pytest.mark.parametrize("C,B,T,beta", [(32, 2, 1000, False), (48, 1, 333, True), (16, 3, 64, False),
                                        (32, 1, 7, False)])
def test_alias_free_snake_and_decoupling_conv(C, B, T, beta):
    """ou_alias_free_snake against the oracle restatement of bigvgan.AliasFreeSnake and of
    UniverseGAN.signal_decoupling_layer (fp32 in, fp32 out: tight tolerance), plus the blocked
    bf16 input flavour enhance() uses."""
    from open_universe_b200.networks.universe.blocks import PReLU_Conv
    torch.manual_seed(5)
    pc = PReLU_Conv(C, 1, kernel_size=3, padding="same", act_type="snakebeta" if beta else "snake")
    with torch.no_grad():
        pc.prelu.act.act.alpha.copy_(0.3 * torch.randn(C))
        if beta:
            pc.prelu.act.act.beta.copy_(0.3 * torch.randn(C))
    sd = {"p." + k: v.detach() for k, v in pc.state_dict().items()}
    x = torch.randn(B, C, T)
    want_act = O.alias_free_snake(sd, "p.prelu", x, beta=beta)
    cfg = {"losses": {"signal_decoupling_act": "snakebeta" if beta else "snake"}}
    sd2 = {k.replace("p.", "signal_decoupling_layer."): v for k, v in sd.items()}
    want_wav = O.aux_to_wav(cfg, sd2, x)
    pc = pc.to(DEV)
    got_act = pc.prelu(x.to(DEV)).cpu()
    got_wav = pc(x.to(DEV)).cpu()
    assert got_act.shape == want_act.shape and got_wav.shape == want_wav.shape == (B, 1, T)
    assert rel_rms(got_act, want_act) < 2e-5
    assert rel_rms(got_wav, want_wav) < 2e-5
    xb = bf(x)
    want_b = O.aux_to_wav(cfg, sd2, xb)
    got_b = R.prelu_conv_forward(pc, R.pack_blocked(xb.to(DEV)), blocked=True).cpu()
    assert rel_rms(got_b, want_b) < 2e-5


def test_conv_resident_weights_with_more_k_blocks_than_a_stages():
    """Regression: a layer whose weights stay resident in shared memory but that has more K blocks
    than A stages (UNIVERSE++ 24 kHz conditioner enc.1.down: 96 -> 192, stride 3, k = 3 -> 9 K blocks,
    8 stages) must not wait for the whole weight slice before releasing A stages (deadlock)."""
    g = torch.Generator().manual_seed(11)
    B, cin, cout, s, t = 2, 96, 192, 3, 3 * 700
    fc = rand_fc(g, cin, cout, s=s, taps=1, tap_off=0, prelu_in=0.25)
    prog = P.Program(B)
    prog.buf("in", "blocked", cin, t)
    P.add_conv(prog, "c", "in", "out", fc, t)
    x = bf(torch.randn(B, cin, t, generator=g))
    bufs = {"in": x}
    E.run_conv(prog.ops[0], bufs, quant=True)
    exe = R.Executor(prog, DEV, external=["in"])
    exe.bufs["in"] = R.pack_blocked(x.to(DEV))
    exe.run()
    torch.cuda.synchronize()
    assert rel_rms(R.unpack_blocked(exe.bufs["out"]).cpu(), bufs["out"]) < 3e-3


TAIL_CASES = [
    dict(t=1000, B=2, sc=True, film=True, skip=True, t_up=2000),
    dict(t=123, B=1, sc=False, film=False, skip=True, t_up=245),       # one partial item, odd target length
    dict(t=122 * 3, B=2, sc=True, film=False, skip=False, t_up=122 * 6),  # exact multiple of the item, no skip
    dict(t=122 * 2 + 1, B=1, sc=True, film=True, skip=True, t_up=2 * (122 * 2 + 1)),
    dict(t=64080, B=3, sc=True, film=True, skip=True, t_up=128160),     # many items per CTA
    dict(t=1000, B=2, sc=True, film=True, skip=True, t_up=2000, taps=1),       # plain k = s up conv (UNIVERSE original)
    dict(t=122 * 2 + 1, B=1, sc=False, film=False, skip=True, t_up=2 * (122 * 2 + 1) - 1, taps=1),
]


@pytest.mark.parametrize("c", TAIL_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_trunk_with_up_tail_vs_emulator(c):
    """ou_conv_trunk with the next block's transposed up conv (+ skip add) fused behind conv3 (C = 64):
    against the emulator and against the four separate launches."""
    g = torch.Generator().manual_seed(987)
    B, C, t, t_up = c["B"], 64, c["t"], c["t_up"]
    prog = P.Program(B)
    prog.buf("in", "blocked", C, t)
    inputs = {"in": bf(torch.randn(B, C, t, generator=g))}
    if c["sc"]:
        prog.buf("sc", "blocked", C, t)
        inputs["sc"] = bf(torch.randn(B, C, t, generator=g))
    if c["skip"]:
        prog.buf("skip", "blocked", C // 2, t_up)
        inputs["skip"] = bf(torch.randn(B, C // 2, t_up, generator=g))
    P.add_conv(prog, "conv1", "in", "c1", rand_fc(g, C, C, taps=5, tap_off=-2, prelu_in=0.2), t,
               add1="sc" if c["sc"] else None, scale1=0.7071 if c["sc"] else 1.0,
               film_off=0 if c["film"] else None, prelu_out=0.15)
    P.add_conv(prog, "conv2", "c1", "c2", rand_fc(g, C, C, taps=3, tap_off=-1), t, prelu_out=0.3)
    P.add_conv(prog, "conv3", "c2", "v", rand_fc(g, C, C, taps=3, tap_off=-1), t, add1="in", scale1=0.7071)
    assert P.fuse_trunk(prog, "trunk")
    taps = c.get("taps", 3)
    P.add_conv(prog, "up", "v", "h", rand_fc(g, C, C // 2, up=2, taps=taps, tap_off=-(taps // 2), prelu_in=0.25), t, t_up,
               add1="skip" if c["skip"] else None, scale1=0.7071 if c["skip"] else 1.0)
    assert P.fuse_up_tail(prog) and len(prog.ops) == 1 and prog.ops[0].tail is not None
    film = torch.randn(B, 2 * C, generator=g) if c["film"] else None
    bufs, _, _ = E.run_program(prog, inputs, film=film, quant=True)
    want = bufs["h"]

    exe = R.Executor(prog, DEV, external=list(inputs))
    for k, v in inputs.items():
        exe.bufs[k] = R.pack_blocked(v.to(DEV))
    film_d = film.to(DEV).contiguous() if film is not None else None
    exe.bufs["h"].fill_(7.0)
    n0 = lib.launch_count()
    exe.run(film=film_d, film_bstride=2 * C)
    assert lib.launch_count() - n0 == 1
    got = R.unpack_blocked(exe.bufs["h"]).cpu()
    R.USE_TRUNK = False
    try:
        exe.bufs["h"].zero_()
        n0 = lib.launch_count()
        exe.run(film=film_d, film_bstride=2 * C)
        assert lib.launch_count() - n0 == 4
        split = R.unpack_blocked(exe.bufs["h"]).cpu()
    finally:
        R.USE_TRUNK = True
    assert got.shape == want.shape and torch.isfinite(got).all()
    err, err_split = rel_rms(got, want), rel_rms(got, split)
    if err >= 3e-3 or err_split >= 3e-3:
        d = (got - want).abs()
        bad = (d > 0.05 * want.abs().max()).float()
        raise AssertionError(f"rel_rms={err:.4f} vs split {err_split:.4f} bad_fraction={bad.mean():.5f} "
                             f"bad rows (time) {bad.amax(dim=(0, 1)).nonzero().flatten()[:24].tolist()}")


OUT_TAIL_CASES = [
    dict(t=1000, B=2, sc=True, film=True, sde=True, noise=True),
    dict(t=251, B=1, sc=False, film=False, sde=False, noise=False),      # raw network output only, partial item
    dict(t=250 * 3, B=3, sc=True, film=False, sde=True, noise=False),    # exact multiple of the item
    dict(t=250 * 2 + 1, B=2, sc=False, film=True, sde=True, noise=True),
    dict(t=128160, B=2, sc=True, film=True, sde=True, noise=True),       # many items per CTA
]


@pytest.mark.parametrize("c", OUT_TAIL_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_trunk_with_output_tail_vs_emulator(c):
    """ou_conv_trunk with the network's output conv + EDM mix + SDE update fused behind conv3 (C = 32):
    against the emulator and against the separate launches (trunk expansion + ou_output_sde)."""
    g = torch.Generator().manual_seed(555)
    B, C, t = c["B"], 32, c["t"]
    prog = P.Program(B)
    prog.buf("in", "blocked", C, t)
    inputs = {"in": bf(torch.randn(B, C, t, generator=g))}
    if c["sc"]:
        prog.buf("sc", "blocked", C, t)
        inputs["sc"] = bf(torch.randn(B, C, t, generator=g))
    P.add_conv(prog, "conv1", "in", "c1", rand_fc(g, C, C, taps=5, tap_off=-2, prelu_in=0.2), t,
               add1="sc" if c["sc"] else None, scale1=0.7071 if c["sc"] else 1.0,
               film_off=0 if c["film"] else None, prelu_out=0.15)
    P.add_conv(prog, "conv2", "c1", "c2", rand_fc(g, C, C, taps=3, tap_off=-1), t, prelu_out=0.3)
    P.add_conv(prog, "conv3", "c2", "v", rand_fc(g, C, C, taps=3, tap_off=-1), t, add1="in", scale1=0.7071,
               prelu_out=0.2)
    assert P.fuse_trunk(prog, "trunk")
    prog.ops.append(P.OutputOp("output_conv", "v", torch.randn(C, 3, generator=g) / 6, 0.3, t, t))
    assert P.fuse_out_tail(prog) and len(prog.ops) == 1 and prog.ops[0].tail_out is not None
    film = torch.randn(B, 2 * C, generator=g) if c["film"] else None
    x = torch.randn(B, 1, t, generator=g)
    coef = torch.randn(B, 3, generator=g) if c["sde"] else None
    noise = torch.randn(B, 1, t, generator=g) if c["noise"] else None
    _, want_net, want_x = E.run_program(prog, dict(inputs, x=x), film=film, coef=coef, noise=noise, quant=True)

    exe = R.Executor(prog, DEV, external=list(inputs))
    for k, v in inputs.items():
        exe.bufs[k] = R.pack_blocked(v.to(DEV))
    exe.bufs["x"] = x.to(DEV)
    film_d = film.to(DEV).contiguous() if film is not None else None
    d_coef = coef.to(DEV) if coef is not None else None
    d_noise = noise.to(DEV) if noise is not None else None

    def run():
        xo = torch.full((B, 1, t), 7.0, device=DEV) if coef is not None else None
        no = torch.full((B, 1, t), 7.0, device=DEV)
        n0 = lib.launch_count()
        exe.run(film=film_d, film_bstride=2 * C, coef=d_coef, noise=d_noise, xout=xo, net_out=no)
        torch.cuda.synchronize()
        return lib.launch_count() - n0, no.cpu(), (xo.cpu() if xo is not None else None)

    n, net, xnew = run()
    assert n == 1
    R.USE_TRUNK = False
    try:
        n, net_split, xnew_split = run()
        assert n == 4
    finally:
        R.USE_TRUNK = True
    assert torch.isfinite(net).all()
    err, err_split = rel_rms(net, want_net), rel_rms(net, net_split)
    if err >= 3e-3 or err_split >= 3e-3:
        bad = ((net - want_net).abs() > 0.05 * want_net.abs().max()).float()
        raise AssertionError(f"net rel_rms={err:.4f} vs split {err_split:.4f} bad_fraction={bad.mean():.5f} "
                             f"bad rows (time) {bad.amax(dim=(0, 1)).nonzero().flatten()[:24].tolist()}")
    if coef is not None:
        assert rel_rms(xnew, want_x) < 3e-3 and rel_rms(xnew, xnew_split) < 3e-3
    # without the SDE arguments the raw output is still written and x_out is left alone


DOWN_TAIL_CASES = [
    dict(t=1000, B=2, film=True),
    dict(t=123, B=1, film=False),             # one partial item
    dict(t=248 * 3, B=2, film=False),         # exact multiple of the item
    dict(t=248 * 2 + 2, B=1, film=True),      # one row pair into a third item
    dict(t=128160, B=2, film=True),           # many items per CTA
    dict(t=1000, B=2, film=True, taps=1),     # plain k = s down conv (UNIVERSE original)
    dict(t=248 * 2 + 3, B=1, film=False, taps=1),
]


@pytest.mark.parametrize("c", DOWN_TAIL_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_conv_trunk_with_down_tail_vs_emulator(c):
    """ou_conv_trunk with the block's own anti-aliased stride-2 down conv (32 -> 64 channels) fused behind conv3:
    block output (skip connection) AND down-conv output against the emulator and against the four separate
    launches."""
    g = torch.Generator().manual_seed(4242)
    B, C, t = c["B"], 32, c["t"]
    prog = P.Program(B)
    prog.buf("in", "blocked", C, t)
    inputs = {"in": bf(torch.randn(B, C, t, generator=g))}
    P.add_conv(prog, "conv1", "in", "c1", rand_fc(g, C, C, taps=5, tap_off=-2, prelu_in=0.2), t,
               film_off=0 if c["film"] else None, prelu_out=0.15)
    P.add_conv(prog, "conv2", "c1", "c2", rand_fc(g, C, C, taps=3, tap_off=-1), t, prelu_out=0.3)
    P.add_conv(prog, "conv3", "c2", "v", rand_fc(g, C, C, taps=3, tap_off=-1), t, add1="in", scale1=0.7071)
    assert P.fuse_trunk(prog, "trunk")
    taps = c.get("taps", 3)
    P.add_conv(prog, "down", "v", "h", rand_fc(g, C, 2 * C, s=2, taps=taps, tap_off=-(taps // 2), prelu_in=0.25), t)
    assert P.fuse_down_tail(prog) and len(prog.ops) == 1 and prog.ops[0].tail_dn is not None
    film = torch.randn(B, 2 * C, generator=g) if c["film"] else None
    bufs, _, _ = E.run_program(prog, inputs, film=film, quant=True)
    want_v, want_h = bufs["v"], bufs["h"]

    exe = R.Executor(prog, DEV, external=list(inputs))
    for k, v in inputs.items():
        exe.bufs[k] = R.pack_blocked(v.to(DEV))
    film_d = film.to(DEV).contiguous() if film is not None else None
    exe.bufs["h"].fill_(7.0)
    exe.bufs["v"].fill_(7.0)
    n0 = lib.launch_count()
    exe.run(film=film_d, film_bstride=2 * C)
    assert lib.launch_count() - n0 == 1
    got_v, got_h = R.unpack_blocked(exe.bufs["v"]).cpu(), R.unpack_blocked(exe.bufs["h"]).cpu()
    R.USE_TRUNK = False
    try:
        exe.bufs["h"].zero_()
        n0 = lib.launch_count()
        exe.run(film=film_d, film_bstride=2 * C)
        assert lib.launch_count() - n0 == 4
        split_h = R.unpack_blocked(exe.bufs["h"]).cpu()
    finally:
        R.USE_TRUNK = True
    assert got_h.shape == want_h.shape and torch.isfinite(got_h).all() and torch.isfinite(got_v).all()
    assert rel_rms(got_v, want_v) < 3e-3
    err, err_split = rel_rms(got_h, want_h), rel_rms(got_h, split_h)
    if err >= 3e-3 or err_split >= 3e-3:
        bad = ((got_h - want_h).abs() > 0.05 * want_h.abs().max()).float()
        raise AssertionError(f"rel_rms={err:.4f} vs split {err_split:.4f} bad_fraction={bad.mean():.5f} "
                             f"bad rows (time) {bad.amax(dim=(0, 1)).nonzero().flatten()[:24].tolist()}")
