"""Precision-policy study on CPU (test infrastructure, uses oracle/): error of ONE ScoreNetwork
evaluation and of a short enhance() under different storage / MMA-operand types, emulated op by op
(oracle/emulator.py rounds exactly where the kernels do).
    python tools/precision_study.py [--seconds 1] [--gain 1]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden"):
    sys.path.insert(0, str(p))
from open_universe_b200.config import builtin_config, instantiate  # noqa: E402
from open_universe_b200.engine import program as P  # noqa: E402
from oracle import emulator as E  # noqa: E402
from oracle.universe_oracle import UniverseOracle, sigma_embedding  # noqa: E402
from emul_enhance import emulated_enhance  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="universepp_16k")
ap.add_argument("--seconds", type=float, default=1.0)
ap.add_argument("--gain", type=float, default=1.0, help="multiply every weight_g of the score net")
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()

torch.manual_seed(1234)
cfg = builtin_config(a.cfg).model
m = instantiate(cfg, _recursive_=False)
m.eval(no_ema=True)
if a.gain != 1.0:
    with torch.no_grad():
        for n, p in m.get_score_model().named_parameters():
            if n.endswith("weight_g"):
                p.mul_(a.gain)
o = UniverseOracle(cfg, m.state_dict())
T = int(m.fs * a.seconds)
g = torch.Generator().manual_seed(3)
mix = 0.05 * torch.randn(1, T, generator=g)
mixp, pad = m.pad(mix[:, None, :])
mixn = o.normalize(mixp)
t_pad = mixn.shape[-1]
noise = [torch.randn(1, 1, t_pad, generator=g) for _ in range(a.steps)]


def rel(x, y):
    return float((x - y).square().mean().sqrt() / y.square().mean().sqrt())


with torch.no_grad():
    cond, _, _ = o.condition(mixn, mixn)
    net = m.get_score_model()
    sp = P.lower_score_network(net, 1, t_pad)
    pp = P.lower_cond_projection(net, 1, sp.meta["lengths"])
    for sig in (2.0, 0.1, 0.005):
        sigma = torch.tensor([sig])
        x = mixn + sig * noise[0]
        want = o.net(x, sigma, cond)
        gs = sigma_embedding(o.cfg["score_model"], o.sd, o.score_prefix + ".sigma_block", torch.log10(sigma))
        film = E.film_table(sp, gs)
        for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
            E.QDTYPE = dt
            pb, _, _ = E.run_program(pp, {f"cond{i}": c for i, c in enumerate(cond)}, quant=True)
            inputs = {"x": x}
            inputs.update({k: v for k, v in pb.items() if k.startswith("sc")})
            _, got, _ = E.run_program(sp, inputs, film=film, quant=True)
            print(f"sigma {sig:6.3f}  {name}: per-evaluation rel rms err {rel(got, want):.3e}   (net rms {float(want.square().mean().sqrt()):.3e})")
    want = o.enhance(mix, n_steps=a.steps, noise=noise)
    zero = None
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        E.QDTYPE = dt
        got = emulated_enhance(m, o, mix, a.steps, noise, True)
        err = float((got - want).square().mean().sqrt())
        print(f"enhance {a.steps} steps {name}: abs rms {err:.3e} rel {rel(got, want):.3e}")
