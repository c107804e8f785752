"""GRU kernel timing vs batch (cluster-wave quantisation check).  GPU box only."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import lib, runtime as R

H, T = 256, 801
L = lib.load()
for B in (4, 8, 16, 24, 28, 32, 48, 64):
    gx = torch.randn(B, T, 6 * H, device="cuda")
    w = torch.randn(2, 3 * H, H, device="cuda") / 16
    b = torch.zeros(2, 3 * H, device="cuda")
    out = R.alloc_blocked(B, 2 * H, T, "cuda")
    for _ in range(2):
        lib.check(L.ou_gru_bidir(R._ptr(gx), R._ptr(w), R._ptr(b), None, 1.0, R._ptr(out), B, T, H, R._stream()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lib.check(L.ou_gru_bidir(R._ptr(gx), R._ptr(w), R._ptr(b), None, 1.0, R._ptr(out), B, T, H, R._stream()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"B={B:3d} clusters={2 * ((B + 3) // 4):3d}  {ms:8.3f} ms  {ms * 1e3 / T:6.2f} us/step")

# per-phase clock64 stamps of the tensor-core kernel (CTA 0, thread 0, steps 100..163)
B = 32
gx = torch.randn(B, T, 6 * H, device="cuda")
out = R.alloc_blocked(B, 2 * H, T, "cuda")
tr = torch.zeros(64, 8, dtype=torch.int64, device="cuda")
L.ou_debug_set_trace(R._ptr(tr))
lib.check(L.ou_gru_bidir(R._ptr(gx), R._ptr(w), R._ptr(b), None, 1.0, R._ptr(out), B, T, H, R._stream()))
torch.cuda.synchronize()
L.ou_debug_set_trace(None)
tr = tr.cpu()
if int(tr.max()) > 0:
    import os
    impl = os.environ.get("OU_GRU_IMPL", "f16")
    if impl.startswith("f1"):
                order = [0, 1, 2, 3, 4, 5]
        names = ["h_full wait", "lds+mma", "gate math", "gather+st.async", "store+mov"]
    else:
        order = [0, 1, 2, 3, 4, 5, 6]
        names = ["h_full wait", "mma phase", "syncthreads", "gate math", "shfl+st.async", "store+mov"]
    tr = tr[:, order]
    d = (tr[:, 1:] - tr[:, :-1]).float().mean(0)
    step = (tr[1:, 0] - tr[:-1, 0]).float().mean()
    back = (tr[1:, 0] - tr[:-1, -1]).float().mean()
    print("GRU per-step cycles (B=32):", {n: round(float(x)) for n, x in zip(names, d)},
          "loop back-edge", round(float(back)), "step", round(float(step)))
