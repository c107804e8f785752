"""PyTorch-eager CUDA timing of the oracle restatement (same op sequence as the reference's own
enhance(): cuDNN convs with TF32 allowed, cuDNN GRU, ATen element-wise) on BASELINE cfg-2.
Reported for context in profiles/README.md -- the reference itself cannot travel to the GPU box.
    python tools/eager_gpu_baseline.py [--batch 32] [--steps 64]"""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.config import builtin_config, instantiate  # noqa: E402
from oracle.universe_oracle import UniverseOracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=8.0)
ap.add_argument("--steps", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
torch.manual_seed(0)
cfg = builtin_config("universepp_16k").model
model = instantiate(cfg, _recursive_=False)
o = UniverseOracle(cfg, model.state_dict()).to("cuda")
mix = 0.05 * torch.randn(a.batch, int(16000 * a.seconds), device="cuda")
print("cudnn.allow_tf32 =", torch.backends.cudnn.allow_tf32, " matmul.allow_tf32 =", torch.backends.cuda.matmul.allow_tf32)
with torch.no_grad():
    for i in range(a.reps + 1):
        rng = torch.Generator(device="cuda").manual_seed(1028282)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        y = o.enhance(mix, n_steps=a.steps, rng=rng)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"run {i}: {dt * 1e3:8.1f} ms  -> {a.batch * a.seconds / dt:8.1f} audio-s/s" + ("  (warm-up)" if i == 0 else ""))
