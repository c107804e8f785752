"""Per-kernel share of ONE enhance() at the bench size from an ncu launch list taken at a reduced
diffusion-step count: the list of `bench.py --diffusion-steps 4` is split, per enhance() call, into the
once-per-call part and the score steps (the launches of the captured sampler loop, 32 per step, right
before `unpad_limit_kernel`), and the score part is scaled to 64 steps.
    python tools/ncu_share.py gpurun_out/launches.csv [steps_in_list=4] [steps_target=64] [launches_per_step=32]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
n_list = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_target = int(sys.argv[3]) if len(sys.argv) > 3 else 64
PER_STEP = int(sys.argv[4]) if len(sys.argv) > 4 else 32
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
launches = OrderedDict()
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
    launches[int(r["ID"])] = (r["Kernel Name"].split("(")[0], us)
seq = [launches[k] for k in sorted(launches)]
ends = [i for i, (n, _) in enumerate(seq) if "unpad_limit" in n]
calls = []
prev = 0
for e in ends:
    calls.append(seq[prev:e + 1])
    prev = e + 1
call = calls[-1]                      # the last (warm) enhance() of the run
score = call[-1 - PER_STEP * n_list:-1]
once = call[:-1 - PER_STEP * n_list] + call[-1:]
ours = lambda n: ("ou::" in n) or n.startswith("tc::") or n.startswith("void trunk::") or n.startswith("void ou::") or "trunk_kernel" in n or "conv1d" in n
tot = OrderedDict()
for part, scale in ((once, 1.0), (score, n_target / n_list)):
    for n, us in part:
        key = n if ours(n) else "(torch element-wise / RNG / copies)"
        tot[key] = tot.get(key, 0.0) + us * scale
total = sum(tot.values())
print(f"one enhance() at {n_target} diffusion steps, predicted from the ncu launch list ({len(call)} launches in the "
      f"listed call, {len(score)} of them score-step launches): {total / 1e3:.1f} ms of kernel time")
for n, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {n[:70]:70s} {us / 1e3:9.2f} ms {100 * us / total:5.1f} %")
conv = sum(us for n, us in tot.items() if "conv1d_tc" in n or "trunk_kernel" in n)
print(f"  conv1d_tc_kernel + trunk_kernel share: {100 * conv / total:.1f} %  (bench.py roofline.share_of_step is the live counterpart)")
