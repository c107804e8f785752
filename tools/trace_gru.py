"""Per-phase cycle counts of one step of the cluster GRU kernel (thread 0 of CTA 0, steps 100..163).
GPU box only:  python tools/trace_gru.py [hidden] [batch] [T]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import lib, runtime as R  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 801
L = lib.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
gx = torch.randn(B, T, 6 * H, device=dev, generator=g)
w_hh = torch.randn(2, 3 * H, H, device=dev, generator=g) / H**0.5
b_hh = torch.zeros(2, 3 * H, device=dev)
out = R.alloc_blocked(B, 2 * H, T, dev)


def run():
    lib.check(L.ou_gru_bidir(R._ptr(gx), R._ptr(w_hh), R._ptr(b_hh), None, 1.0, R._ptr(out), B, T, H, R._stream()))


run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 5 * 1e3
tr = torch.zeros(1024, dtype=torch.int64, device=dev)
L.ou_debug_set_trace(R._ptr(tr))
run()
torch.cuda.synchronize()
L.ou_debug_set_trace(None)
t = tr[:512].reshape(64, 8)[:, :6].cpu().double()
names = ["wait h", "mma", "gates", "shuffle+push", "store+rotate", "loop top (prefetch, arm)"]
d = torch.cat([t[:, 1:] - t[:, :-1], (t[1:, 0:1] - t[:-1, 5:6]).mean(0, keepdim=True).expand(64, 1)], 1)
print(f"H={H} B={B} T={T}: {us:.1f} us = {us / T * 1e3:.0f} ns/step; cycles per step {float((t[1:, 0] - t[:-1, 0]).mean()):.0f}")
print("  " + "  ".join(f"{n}={float(d[:, i].mean()):.0f}" for i, n in enumerate(names)))
