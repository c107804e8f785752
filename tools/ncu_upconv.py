"""Launch the two bandwidth-shaped up convs of BASELINE cfg-2 (dec.4.up 64->32 x2, dec.3.up 128->64 x4, with
skip add) for an ncu source-level capture:
    ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -s 2 -c 2 -o gpurun_out/prof_up \
        python tools/ncu_upconv.py"""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import program as P, runtime as R  # noqa: E402
from open_universe_b200.engine.fold import FoldedConv  # noqa: E402

B = 32
CASES = [("dec.4.up 64->32 x2", 64, 32, 1, 2, 3, 64080, 0.25, True),
         ("dec.3.up 128->64 x4", 128, 64, 1, 4, 3, 16020, 0.25, True)]
g = torch.Generator().manual_seed(0)
exes = []
for name, c, cout, s_, up, taps, t, prelu, add1 in CASES:
    fc = FoldedConv(torch.randn(up * cout, taps, s_ * c, generator=g) / math.sqrt(taps * c * s_),
                    torch.zeros(up * cout), c, cout, s_, up, taps, -(taps // 2), prelu)
    prog = P.Program(B)
    prog.buf("in", "blocked", c, t)
    _, t_out = P.add_conv(prog, "c", "in", "out", fc, t)
    prog.buf("add1", "blocked", cout, t_out)
    prog.ops[0].add1, prog.ops[0].scale1 = "add1", 0.7071
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    exe.bufs["add1"].normal_()
    exes.append((name, exe))
for rep in range(2):
    for name, exe in exes:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        exe.run()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) * 1e3:.1f} us")
