"""Timeline of the tcgen05 conv kernel's warp roles (CTA 0) from clock64 stamps.  GPU box only."""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import lib, program as P, runtime as R
from open_universe_b200.engine.fold import FoldedConv

B = 32
CASES = [("C32 k3", 32, 3, 128160, None, False), ("C32 k3 add", 32, 3, 128160, None, True),
         ("C128 k5 prelu", 128, 5, 16020, 0.25, False), ("C512 k5 prelu", 512, 5, 801, 0.25, False)]
L = lib.load()
g = torch.Generator().manual_seed(0)
for name, c, taps, t, prelu, add1 in CASES:
    fc = FoldedConv(torch.randn(c, taps, c, generator=g) / math.sqrt(taps * c), torch.zeros(c), c, c, 1, 1,
                    taps, -(taps // 2), prelu)
    prog = P.Program(B)
    prog.buf("in", "blocked", c, t)
    P.add_conv(prog, "c", "in", "out", fc, t)
    if add1:
        prog.buf("add1", "blocked", c, t)
        prog.ops[0].add1 = "add1"
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    if add1:
        exe.bufs["add1"].normal_()
    exe.run()
    tr = torch.zeros(4, 64, 4, dtype=torch.int64, device="cuda")
    L.ou_debug_set_trace(R._ptr(tr))
    exe.run()
    torch.cuda.synchronize()
    L.ou_debug_set_trace(None)
    tr = tr.cpu()
    t0 = int(tr[tr > 0].min())
    print(f"== {name}: cycles relative to first stamp (CTA 0)")
    print("tile | P:empty_ok issued | X:full0 fullN ready | M:tmem_ok ready0 readyN commit | E:wait full ldone stored")
    for i in range(12):
        row = []
        for role, nev in ((0, 2), (2, 3), (1, 4), (3, 4)):
            row.append(" ".join(f"{int(tr[role, i, e]) - t0:8d}" if tr[role, i, e] > 0 else "       -" for e in range(nev)))
        print(f"{i:4d} | " + " | ".join(row))
