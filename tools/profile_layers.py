"""Per-op device timing of one score-network evaluation / conditioner pass (CUDA events around
every launch).  Usage (on the GPU box):  python tools/profile_layers.py [--batch 32] [--seconds 8]
Writes a table to stdout; bench-grade numbers come from bench.py, this is the optimisation map."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.config import builtin_config, instantiate  # noqa: E402
from open_universe_b200.engine import program as P, runtime as R, lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--config", default="universepp_16k")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = instantiate(builtin_config(a.config).model, _recursive_=False)
    m.eval(no_ema=True)
    m = m.to(dev)
    T = int(m.fs * a.seconds)
    t_pad = T + (m.tot_ds - T % m.tot_ds)
    B = a.batch
    x = torch.randn(B, 1, t_pad, device=dev) * 0.05
    cr = R.get_conditioner_runner(m.condition_model, B, t_pad, dev, False)
    sr = R.get_score_runner(m.get_score_model(), B, t_pad, dev)

    def timed_ops(exe, run):
        recs = None
        for _ in range(a.reps):
            evs = []
            orig = {}
            L = lib.load()
            # wrap every op launch with events by running ops one at a time
            ops = exe.prog.ops
            full = exe.prog.ops
            for op in ops:
                exe.prog.ops = [op]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run()
                e1.record()
                evs.append((op, e0, e1))
            exe.prog.ops = full
            torch.cuda.synchronize()
            cur = [(op, e0.elapsed_time(e1) * 1e3) for op, e0, e1 in evs]
            recs = cur if recs is None else [(o, min(t0, t1)) for (o, t0), (_, t1) in zip(recs, cur)]
        return recs

    cond, _, _ = cr.run(x, x)
    sr.set_sigmas(torch.full((1,), 0.5, device=dev))
    sr.set_cond(cond)
    coef = torch.tensor([[1.0, 0.1, 0.0]], device=dev).expand(B, 3).contiguous()
    in_scale = torch.ones(B, device=dev)
    xo = torch.empty_like(x)

    def run_score():
        sr.exe.bufs["x"] = x
        sr.exe.run(film=sr.film, film_bstride=0, in_scale=in_scale, coef=coef, xout=xo)

    def run_cond():
        cr.exe.bufs["x"] = x
        cr.exe.bufs["x_wav"] = x
        cr.exe.run()

    for title, exe, run in (("score step", sr.exe, run_score), ("conditioner", cr.exe, run_cond)):
        run()
        recs = timed_ops(exe, run)
        tot = sum(t for _, t in recs)
        print(f"== {title}: {len(recs)} launches, {tot / 1e3:.3f} ms total (B={B}, T={t_pad})")
        print(f"{'op':28s} {'kind':9s} {'cin':>5s} {'n':>5s} {'taps':>4s} {'s':>3s} {'up':>3s} {'rows':>7s} "
              f"{'us':>9s} {'TF/s':>7s} {'GB/s':>7s} {'%':>5s}")
        for op, t in recs:
            kind = type(op).__name__
            if isinstance(op, P.ConvOp):
                fc = op.fc
                byts = 2.0 * B * (fc.cin * op.t_in + fc.cout * op.t_out * (1 + (op.add1 is not None) + (op.add2 is not None)))
                if op.dst_kind != "blocked":
                    byts = 2.0 * B * fc.cin * op.t_in + 4.0 * B * op.rows * fc.n
                print(f"{op.name:28s} {kind:9s} {fc.cin:5d} {fc.n:5d} {fc.taps:4d} {fc.s:3d} {fc.up:3d} {op.rows:7d} "
                      f"{t:9.1f} {op.flops_exec / t / 1e6:7.1f} {byts / t / 1e3:7.0f} {100 * t / tot:5.1f}")
            elif isinstance(op, P.TrunkOp):
                c1 = op.parts[0]
                c = c1.fc.cin
                byts = 2.0 * B * c * c1.t_in * (2 + (c1.add1 is not None))
                print(f"{op.name:28s} {kind:9s} {c:5d} {c:5d} {'533':>4s} {1:3d} {1:3d} {c1.rows:7d} "
                      f"{t:9.1f} {op.flops_exec / t / 1e6:7.1f} {byts / t / 1e3:7.0f} {100 * t / tot:5.1f}")
            else:
                print(f"{op.name:28s} {kind:9s} {'':5s} {'':5s} {'':4s} {'':3s} {'':3s} {'':7s} {t:9.1f} {'':7s} {'':7s} {100 * t / tot:5.1f}")


if __name__ == "__main__":
    main()
