"""Launch ONE instance of each hot kernel of BASELINE cfg-2 (B=32, 8 s clips) for an `ncu --set full`
capture (profiles/README.md lists the command).  GPU box only; also prints CUDA-event timings."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import lib, program as P, runtime as R  # noqa: E402
from open_universe_b200.engine.fold import FoldedConv  # noqa: E402

B = 32
g = torch.Generator().manual_seed(0)
L = lib.load()


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


def conv(name, cin, cout, s, up, taps, t, prelu, add1):
    fc = FoldedConv(torch.randn(up * cout, taps, s * cin, generator=g) / math.sqrt(taps * cin * s),
                    torch.zeros(up * cout), cin, cout, s, up, taps, -(taps // 2), prelu)
    prog = P.Program(B)
    prog.buf("in", "blocked", cin, t)
    _, t_out = P.add_conv(prog, "c", "in", "out", fc, t)
    if add1:
        prog.buf("add1", "blocked", cout, t_out)
        prog.ops[0].add1, prog.ops[0].scale1 = "add1", 0.7071
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    if add1:
        exe.bufs["add1"].normal_()
    us = timed(exe.run)
    op = prog.ops[0]
    byts = 2.0 * B * (cin * t + cout * t_out * (2 if add1 else 1))
    print(f"{name:34s} {us:8.1f} us  {op.flops_exec / us / 1e6:7.1f} TFLOP/s executed  {byts / us / 1e3:6.0f} GB/s algorithmic")


def trunk(name, c, t, sc, tail=None):
    """tail: None | 'up' (C = 64: next block's x2 up conv + skip) | 'down' (C = 32: own stride-2 down conv) |
    'out' (C = 32: output conv + EDM / SDE update)."""
    def fc(taps, prelu):
        return FoldedConv(torch.randn(c, taps, c, generator=g) / math.sqrt(taps * c), torch.zeros(c), c, c, 1, 1,
                          taps, -(taps // 2), prelu)
    prog = P.Program(B)
    prog.buf("in", "blocked", c, t)
    if sc:
        prog.buf("sc", "blocked", c, t)
    P.add_conv(prog, "conv1", "in", "c1", fc(5, 0.25), t, add1="sc" if sc else None, scale1=0.7071,
               film_off=0, prelu_out=0.25)
    P.add_conv(prog, "conv2", "c1", "c2", fc(3, None), t, prelu_out=0.25)
    P.add_conv(prog, "conv3", "c2", "v", fc(3, None), t, add1="in", scale1=0.7071)
    assert P.fuse_trunk(prog, "trunk")
    kw = {}
    if tail == "up":
        prog.buf("skip", "blocked", c // 2, 2 * t)
        fu = FoldedConv(torch.randn(2 * (c // 2), 3, c, generator=g) / math.sqrt(3 * c), torch.zeros(c), c, c // 2, 1, 2,
                        3, -1, 0.25)
        P.add_conv(prog, "up", "v", "h", fu, t, 2 * t, add1="skip", scale1=0.7071)
        assert P.fuse_up_tail(prog)
    elif tail == "down":
        fd = FoldedConv(torch.randn(2 * c, 3, 2 * c, generator=g) / math.sqrt(6 * c), torch.zeros(2 * c), c, 2 * c, 2, 1,
                        3, -1, 0.25)
        P.add_conv(prog, "down", "v", "h", fd, t)
        assert P.fuse_down_tail(prog)
    elif tail == "out":
        prog.ops.append(P.OutputOp("output_conv", "v", torch.randn(c, 3, generator=g) / 10, 0.0, t, t))
        assert P.fuse_out_tail(prog)
        kw = dict(coef=torch.randn(B, 3, device="cuda"), noise=torch.randn(B, 1, t, device="cuda"),
                  xout=torch.empty(B, 1, t, device="cuda"))
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    for k in ("sc", "skip"):
        if k in exe.bufs and exe.bufs[k] is not None:
            exe.bufs[k].normal_()
    if tail == "out":
        exe.bufs["x"] = torch.randn(B, 1, t, device="cuda")
    film = torch.randn(1, 2 * c, device="cuda")
    us = timed(lambda: exe.run(film=film, film_bstride=0, **kw))
    byts = P.op_bytes(prog.ops[0], B)
    print(f"{name:34s} {us:8.1f} us  {prog.ops[0].flops_exec / us / 1e6:7.1f} TFLOP/s executed  {byts / us / 1e3:6.0f} GB/s algorithmic")


def gru(H=256, T=801):
    gx = torch.randn(B, T, 6 * H, device="cuda")
    w = torch.randn(2, 3 * H, H, device="cuda") / 16
    b = torch.zeros(2, 3 * H, device="cuda")
    add = R.alloc_blocked(B, 2 * H, T, "cuda").normal_()
    out = R.alloc_blocked(B, 2 * H, T, "cuda")
    us = timed(lambda: lib.check(L.ou_gru_bidir(R._ptr(gx), R._ptr(w), R._ptr(b), R._ptr(add), 0.7071, R._ptr(out),
                                                B, T, H, R._stream())))
    print(f"{'bottleneck BiGRU H=256 T=801':34s} {us:8.1f} us  {us * 1e3 / T:6.0f} ns per recurrence step")


conv("L2 conv1 C128 k5 prelu", 128, 128, 1, 1, 5, 16020, 0.25, False)
conv("L3 conv1 C256 k5 prelu", 256, 256, 1, 1, 5, 4005, 0.25, False)
conv("dec.3.up 128->64 x4 +skip", 128, 64, 1, 4, 3, 16020, 0.25, True)
trunk("enc.1 trunk C64", 64, 64080, False)
trunk("dec.3 trunk C64 +sc +up tail", 64, 64080, True, "up")
trunk("enc.0 trunk C32 +down tail", 32, 128160, False, "down")
trunk("dec.4 trunk C32 +sc +out tail", 32, 128160, True, "out")
gru()
