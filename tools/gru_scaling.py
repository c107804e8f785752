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
