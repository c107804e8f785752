"""Timeline of the two half-batch op streams of the pipelined sampler loop (GPU box only): runs enhance()
kernel by kernel (no graph) with CUDA events around every launch on both streams and prints, for one
diffusion step in the middle, when each op of each half started / ended relative to the step's first op.
    OU_PIPE_PROFILE=1 python tools/pipe_timeline.py [--steps 8] [--show 5]"""
import argparse
import os
import sys
from pathlib import Path

os.environ["OU_PIPE_PROFILE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.config import builtin_config, instantiate  # noqa: E402
from open_universe_b200.engine import program as P, runtime  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--show", type=int, default=5)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=8.0)
ap.add_argument("--config", default="universepp_16k")
a = ap.parse_args()
torch.manual_seed(0)
m = instantiate(builtin_config(a.config).model, _recursive_=False)
m.eval(no_ema=True)
m = m.to("cuda")
x = 0.05 * torch.randn(a.batch, int(m.fs * a.seconds), device="cuda")
rng = torch.Generator(device="cuda").manual_seed(1)
runtime.PROFILE = []
m.enhance(x, n_steps=a.steps, rng=rng)          # warm-up (lazy kernel attributes), also un-captured
torch.cuda.synchronize()
runtime.PROFILE = []
m.enhance(x, n_steps=a.steps, rng=rng)
torch.cuda.synchronize()
prof, runtime.PROFILE = runtime.PROFILE, None
# score-step launches: ops of the two half programs, in issue order; a step starts at each InputConvOp of half A
recs = [(op, b, e0, e1) for op, b, e0, e1 in prof]
starts = [i for i, (op, b, _, _) in enumerate(recs) if isinstance(op, P.InputConvOp) and op.name == "input_conv"
          and op.use_in_scale]
half_a = recs[starts[0]][1]
steps = [i for i in starts if recs[i][1] == half_a and (i == starts[0] or True)]
# keep only the first input conv of each step (half A is issued first)
first = []
for i in starts:
    if not first or i - first[-1] > 40:
        first.append(i)
lo = first[a.show]
hi = first[a.show + 1] if a.show + 1 < len(first) else len(recs)
t0 = recs[lo][2]
print(f"step {a.show}: {hi - lo} launches; times in us relative to the step's first launch")
seen = {}
for op, b, e0, e1 in recs[lo:hi]:
    key = (op.name, b)
    half = "A" if seen.setdefault(op.name, b) == b and list(seen.keys()).count(op.name) == 1 and seen[op.name] == b else "?"
    print(f"B={b:2d} {op.name:16s} {t0.elapsed_time(e0) * 1e3:9.1f} -> {t0.elapsed_time(e1) * 1e3:9.1f}  ({e0.elapsed_time(e1) * 1e3:7.1f})")
