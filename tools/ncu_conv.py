"""Launch a few representative conv layers of BASELINE cfg-2 (B=32, 8 s) for ncu captures.
    ncu --set full --clock-control none --import-source on -k regex:conv1d -c 12 -o gpurun_out/prof \
        python tools/ncu_conv.py
Also prints CUDA-event timings when run without a profiler."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from open_universe_b200.engine import program as P, runtime as R  # noqa: E402
from open_universe_b200.engine.fold import FoldedConv  # noqa: E402

B = 32
LAYERS = [
    # name, cin, cout, taps, t, prelu_in, add1, film
    ("L0.conv2 C32 k3", 32, 32, 3, 128160, None, False, False),
    ("L0.conv3 C32 k3 +add", 32, 32, 3, 128160, None, True, False),
    ("L1.conv1 C64 k5 prelu film", 64, 64, 5, 64080, 0.25, False, True),
    ("L2.conv1 C128 k5 prelu", 128, 128, 5, 16020, 0.25, False, False),
    ("L3.conv1 C256 k5 prelu", 256, 256, 5, 4005, 0.25, False, False),
    ("L4.conv1 C512 k5 prelu", 512, 512, 5, 801, 0.25, False, False),
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    g = torch.Generator().manual_seed(0)
    for name, cin, cout, taps, t, prelu_in, add1, film in LAYERS:
        fc = FoldedConv(torch.randn(cout, taps, cin, generator=g) / math.sqrt(taps * cin),
                        torch.zeros(cout), cin, cout, 1, 1, taps, -(taps // 2), prelu_in)
        prog = P.Program(B)
        prog.buf("in", "blocked", cin, t)
        P.add_conv(prog, "c", "in", "out", fc, t)
        op = prog.ops[0]
        if add1:
            prog.buf("add1", "blocked", cout, t)
            op.add1, op.scale1 = "add1", 0.7071
        filmt = None
        if film:
            op.film_off = 0
            filmt = torch.randn(1, 2 * cout, device="cuda")
        exe = R.Executor(prog, "cuda")
        exe.bufs["in"].copy_(torch.randn(exe.bufs["in"].shape, device="cuda") * 0.5)
        if add1:
            exe.bufs["add1"].normal_()
        times = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            exe.run(film=filmt, film_bstride=0)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) * 1e3)
        us = min(times)
        byts = 2.0 * B * t * (cin + cout * (2 if add1 else 1))
        print(f"{name:30s} {us:9.1f} us  {op.flops_exec / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.0f} GB/s")


if __name__ == "__main__":
    main()
