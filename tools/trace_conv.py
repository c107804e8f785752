"""Timeline of the tcgen05 conv kernel's warp roles (CTA 0) from clock64 stamps.  GPU box only."""
import math
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import lib, program as P, runtime as R
from open_universe_b200.engine.fold import FoldedConv

B = 32
# name, cin, cout, s, up, taps, t_in, prelu_in, add1
CASES = [("enc.0.down 32->64 s2", 32, 64, 2, 1, 3, 128160, 0.25, False),
         ("enc.1.down 64->128 s4", 64, 128, 4, 1, 3, 64080, 0.25, False),
         ("dec.3.up 128->64 x4", 128, 64, 1, 4, 3, 16020, 0.25, True),
         ("dec.4.up 64->32 x2", 64, 32, 1, 2, 3, 64080, 0.25, True),
         ("C128 k5 prelu", 128, 128, 1, 1, 5, 16020, 0.25, False),
         ("C128 k3 add", 128, 128, 1, 1, 3, 16020, None, True),
         ("C128 k3", 128, 128, 1, 1, 3, 16020, None, False)]
L = lib.load()
g = torch.Generator().manual_seed(0)
for name, c, cout, s_, up, taps, t, prelu, add1 in CASES:
    fc = FoldedConv(torch.randn(up * cout, taps, s_ * c, generator=g) / math.sqrt(taps * c * s_), torch.zeros(up * cout),
                    c, cout, s_, up, taps, -(taps // 2), prelu)
    prog = P.Program(B)
    prog.buf("in", "blocked", c, t)
    _, t_out = P.add_conv(prog, "c", "in", "out", fc, t)
    if add1:
        prog.buf("add1", "blocked", cout, t_out)
        prog.ops[0].add1 = "add1"
    exe = R.Executor(prog, "cuda")
    exe.bufs["in"].normal_()
    if add1:
        exe.bufs["add1"].normal_()
    exe.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        exe.run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    trbuf = torch.zeros(2048, dtype=torch.int64, device="cuda")
    tr = trbuf[:1024].view(4, 64, 4)
    L.ou_debug_set_trace(R._ptr(trbuf))
    exe.run()
    torch.cuda.synchronize()
    L.ou_debug_set_trace(None)
    det = trbuf[1024:1024 + 32].cpu().view(8, 4)
    tr = tr.cpu()
    t0 = int(tr[tr > 0].min())
    per = [float((tr[r, 5:12, e] - tr[r, 4:11, e]).float().mean()) for r, e in ((0, 1), (2, 2), (1, 3), (3, 3))]
    print(f"== {name}: {us:.1f} us; steady-state cycles per tile: producer {per[0]:.0f} transform {per[1]:.0f} "
          f"mma {per[2]:.0f} epilogue {per[3]:.0f}; cycles relative to first stamp (CTA 0)")
    if int(det.max()) > 0:
        d0 = int(det[det > 0].min())
        print("   epilogue items of tile 8 (warp 0 of the epilogue, cycles): start, tmem loaded, residual prefetch issued, stored")
        for i in range(8):
            print("   item", i, " ".join(f"{int(v) - d0:7d}" if v > 0 else "      -" for v in det[i]))
    print("tile | P:empty_ok issued | X:full0 fullN ready | M:tmem_ok ready0 readyN commit | E:wait full ldone stored")
    for i in range(3, 8):
        row = []
        for role, nev in ((0, 2), (2, 3), (1, 4), (3, 4)):
            row.append(" ".join(f"{int(tr[role, i, e]) - t0:8d}" if tr[role, i, e] > 0 else "       -" for e in range(nev)))
        print(f"{i:4d} | " + " | ".join(row))
