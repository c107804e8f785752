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


def trunk(name, c, t, sc):
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
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    if sc:
        exe.bufs["sc"].normal_()
    film = torch.randn(1, 2 * c, device="cuda")
    us = timed(lambda: exe.run(film=film, film_bstride=0))
    byts = 2.0 * B * c * t * (3 if sc else 2)
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
conv("dec.4.up 64->32 x2 +skip", 64, 32, 1, 2, 3, 64080, 0.25, True)
trunk("enc.1 trunk C64", 64, 64080, False)
trunk("dec.4 trunk C32 +sc", 32, 128160, True)
gru()
