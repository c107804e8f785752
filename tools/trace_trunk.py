"""Per-stage timeline of the fused trunk kernel (CTA 0: slot-0 warpgroup and the MMA warp) from clock64
stamps.  GPU box only."""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import lib, program as P, runtime as R
from open_universe_b200.engine.fold import FoldedConv

B = 32
L = lib.load()
g = torch.Generator().manual_seed(0)
for name, c, t, sc in (("C64 enc", 64, 64080, False), ("C64 dec", 64, 64080, True), ("C32 enc", 32, 128160, False),
                       ("C32 dec", 32, 128160, True), ("C64 dec + up tail", 64, 64080, True)):
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
    tail = name.endswith("tail")
    if tail:
        prog.buf("skip", "blocked", c // 2, 2 * t)
        fu = FoldedConv(torch.randn(c, 3, c, generator=g) / math.sqrt(3 * c), torch.zeros(c), c, c // 2, 1, 2, 3, -1, 0.25)
        P.add_conv(prog, "up", "v", "h", fu, t, 2 * t, add1="skip", scale1=0.7071)
        assert P.fuse_up_tail(prog)
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    if sc:
        exe.bufs["sc"].normal_()
    film = torch.randn(1, 2 * c, device="cuda")
    exe.run(film=film, film_bstride=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        exe.run(film=film, film_bstride=0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    byts = 2.0 * B * c * t * (3 if sc else 2)
    tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
    L.ou_debug_set_trace(R._ptr(tr))
    exe.run(film=film, film_bstride=0)
    torch.cuda.synchronize()
    L.ou_debug_set_trace(None)
    tr = tr.cpu()
    sl = tr[:512].reshape(32, 16)[:, :16 if tail else 15]
    ok = sl[:, 0] > 0
    sl = sl[ok].double()
    names = ["wait x", "T0", "bar", "issue1", "wait acc1", "E1", "bar", "issue2", "wait acc2", "E2", "bar", "issue3",
             "wait acc3", "E3"] + (["tail: bar+issue+wait+E4"] if tail else [])
    d = sl[:, 1:] - sl[:, :-1]
    per_item = float((sl[1:, 0] - sl[:-1, 0]).float().mean())
    print(f"== {name}: {us:.1f} us, {byts / us / 1e3:.0f} GB/s; slot-0 item period {per_item:.0f} cycles")
    print("   " + "  ".join(f"{n}={float(d[:, i].float().mean()):.0f}" for i, n in enumerate(names)))
