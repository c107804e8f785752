"""Device-timed enhance() of the other BASELINE.json configs (parity-test cases, not bench lines):
cfg-3 UNIVERSE (original) 16 kHz 64 x 4 s x 32 steps; cfg-4 UNIVERSE++ 24 kHz, the 4-clip share of one
of its 4 GPUs (16 x 10 s x 64 steps batch-sharded); cfg-1 the 1 x 2 s x 8 steps smoke case.  GPU box only."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.config import builtin_config, instantiate  # noqa: E402

CASES = [("cfg-1 U++16k 1x2s x8", "universepp_16k", 1, 2.0, 8),
         ("cfg-3 UNIVERSE16k 64x4s x32", "universe_original_16k", 64, 4.0, 32),
         ("cfg-4 U++24k 4x10s x64 (per GPU)", "universepp_24k", 4, 10.0, 64)]
for name, cfg, B, sec, steps in CASES:
    torch.manual_seed(0)
    m = instantiate(builtin_config(cfg).model, _recursive_=False)
    m.eval(no_ema=True)
    m = m.to("cuda")
    x = 0.05 * torch.randn(B, int(m.fs * sec), device="cuda")
    rng = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(2):
        y = m.enhance(x, n_steps=steps, rng=rng)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y = m.enhance(x, n_steps=steps, rng=rng)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    assert torch.isfinite(y).all()
    print(f"{name:36s} {ms:9.2f} ms per enhance()  {B * sec / ms * 1e3:9.1f} audio-s/s")
    del m
    torch.cuda.empty_cache()
