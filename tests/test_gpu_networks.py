"""Network-level and end-to-end parity of the CUDA path (through the drop-in module API, i.e.
through the C ABI) against the golden vectors of the unmodified reference and the CPU oracle.

Precision policy under test: fp16 storage + fp16 tensor-core operands (csrc/common.cuh; bf16 with
OU_ACT_BF16=1), fp32 accumulation, fp32 FiLM / GRU state / EDM + SDE update.  Two weight sets:
  * "det"  -- tests/golden/detweights.py, unit-gain random weights: a deliberately harsh
    stress set (a single score evaluation amplifies a 2^-9 bf16 rounding to ~2e-2 relative
    even when ONLY the MMA operands are rounded, see DESIGN.md "Precision"); tolerances for
    it are relative and loose, its job is to catch wiring / indexing errors (which show up
    as O(1) errors) against the golden files of the real reference;
  * "init" -- the reference's own initialisation scheme (our constructors, seeded): the
    regime of SURVEY.md section 7 / BASELINE.md 3c.  Gate: north_star's 1e-3 RMS (absolute).
"""
import pytest
import torch

from cases import ENHANCE_CASES, NET_CASES, case_kwargs, noise_rows
from common import (OUR_CONFIG, abs_rms, det_audio, det_noise, full_state_dict, load_golden,
                    make_oracle, rel_rms, sub)

pytestmark = pytest.mark.gpu
DEV = "cuda"
_models = {}


def our_model(name):
    if name not in _models:
        from open_universe_b200.config import builtin_config, instantiate
        m = instantiate(builtin_config(OUR_CONFIG[name]).model, _recursive_=False)
        m.load_state_dict(full_state_dict(name), strict=True)
        m.eval(no_ema=True)
        _models[name] = m.to(DEV)
    return _models[name]


def inject_noise(monkeypatch, noise):
    from open_universe_b200.networks.universe import universe as U
    it = iter(noise)

    def randn(x, sigma, rng=None):
        n = next(it).to(x)
        assert n.shape == x.shape
        return n * sigma[:, None, None]

    monkeypatch.setattr(U, "randn", randn)


@pytest.mark.parametrize("case", NET_CASES, ids=lambda c: c["name"])
def test_networks_vs_reference_golden(case):
    g = load_golden(case["name"])
    m = our_model(case["model"])
    o = make_oracle(case["model"])
    B, T = case["B"], case["T"]
    x_wav = det_audio((B, 1, T), case["seed"], level=0.05)
    x_t = det_noise(1, (B, 1, T), case["seed"])[0] * 0.3
    sigma = torch.tensor((case["sigmas"] * B)[:B], dtype=torch.float32)
    mel = m.condition_model.input_mel.compute_mel_spec(x_wav.to(DEV)).cpu()
    assert mel.shape == g["mel"].shape            # same integer frame indexing
    assert rel_rms(mel, g["mel"]) < 2e-5
    cond, y_hat, h = m.condition_model(x_wav.to(DEV), x_wav=x_wav.to(DEV), train=True)
    for name, t in [("y_hat", y_hat), ("h", h)] + [(f"cond{i}", c) for i, c in enumerate(cond)]:
        assert list(t.shape) == list(g[name + "_shape"]), name
        err = rel_rms(sub(t.cpu(), g[name + "_stride"]), g[name])
        assert err < 2e-2, (name, err)
    # score network on the reference's conditioning (isolates it from conditioner error)
    with torch.no_grad():
        cond_ref, _, _ = o.condition(x_wav, x_wav)
    cond_ref = [c.to(DEV) for c in cond_ref]
    score = m.score_model(x_t.to(DEV), sigma.to(DEV), cond_ref).cpu()
    assert score.shape == g["score"].shape
    err = rel_rms(score, g["score"])
    assert err < 6e-2, err


@pytest.mark.parametrize("case", NET_CASES[:2], ids=lambda c: c["name"])
def test_score_network_raw_output(case):
    """ScoreNetwork.forward (no EDM wrapper) against the reference's raw network output."""
    g = load_golden(case["name"])
    m = our_model(case["model"])
    o = make_oracle(case["model"])
    B, T = case["B"], case["T"]
    x_wav = det_audio((B, 1, T), case["seed"], level=0.05)
    x_t = det_noise(1, (B, 1, T), case["seed"])[0] * 0.3
    sigma = torch.tensor((case["sigmas"] * B)[:B], dtype=torch.float32)
    with torch.no_grad():
        cond_ref, _, _ = o.condition(x_wav, x_wav)
    net = m.get_score_model()(x_t.to(DEV), sigma.to(DEV), [c.to(DEV) for c in cond_ref]).cpu()
    assert rel_rms(net, g["net"]) < 6e-2


@pytest.mark.parametrize("case", ENHANCE_CASES, ids=lambda c: c["name"])
def test_enhance_vs_reference_golden(case, monkeypatch):
    g = load_golden(case["name"])
    m = our_model(case["model"])
    shape = tuple(case["shape"])
    mix = det_audio(shape, case["seed"])
    noise = det_noise(case["n_steps"], (noise_rows(case), 1, int(g["t_pad"])), case["seed"])
    inject_noise(monkeypatch, noise)
    kw = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in case_kwargs(case).items()}
    y = m.enhance(mix.to(DEV), n_steps=case["n_steps"], **kw).cpu()
    assert y.shape == mix.shape == g["y"].shape
    assert torch.isfinite(y).all()
    err = rel_rms(y, g["y"])
    print(case["name"], "rel", err, "abs", abs_rms(y, g["y"]))
    assert err < 0.12, err      # stress weights, see module docstring


@pytest.mark.parametrize("cfg_name,shape,n_steps", [
    ("universepp_16k", (1, 32000), 8),          # BASELINE.json configs[0]
    ("universepp_16k", (2, 16000), 16),
    ("universe_original_16k", (1, 16000), 8),
    ("universepp_24k", (1, 12000), 4),
])
def test_enhance_north_star_tolerance(cfg_name, shape, n_steps, monkeypatch, parity_log):
    """north_star gate: CUDA enhance() vs the oracle on identical inputs and identical injected
    diffusion noise, weights drawn by the reference's own init scheme: <= 1e-3 RMS."""
    from open_universe_b200.config import builtin_config, instantiate
    from oracle.universe_oracle import UniverseOracle
    torch.manual_seed(1234)
    cfg = builtin_config(cfg_name).model
    m = instantiate(cfg, _recursive_=False)
    m.eval(no_ema=True)
    o = UniverseOracle(cfg, m.state_dict())
    mix = det_audio(shape, 77)
    t_pad = shape[-1] + (m.tot_ds - shape[-1] % m.tot_ds)
    noise = det_noise(n_steps, (shape[0], 1, t_pad), 77)
    with torch.no_grad():
        want = o.enhance(mix, n_steps=n_steps, noise=noise)
    inject_noise(monkeypatch, noise)
    got = m.to(DEV).enhance(mix.to(DEV), n_steps=n_steps).cpu()
    a, r = abs_rms(got, want), rel_rms(got, want)
    print(cfg_name, shape, n_steps, "abs rms err", a, "rel", r, "out rms", float(want.square().mean().sqrt()))
    parity_log[f"north_star_{cfg_name}_{shape[0]}x{shape[1]}_{n_steps}"] = {
        "config": cfg_name, "shape": list(shape), "n_steps": n_steps, "abs_rms_err": a, "rel_rms_err": r,
        "out_rms": float(want.square().mean().sqrt())}
    assert got.shape == want.shape
    # north_star: 1e-3 RMS.  Measured with fp16 storage (profiles/parity_r2.json): UNIVERSE++ 4e-6 absolute /
    # 1e-4 .. 2.5e-4 relative; UNIVERSE (no EDM wrapper, output RMS 0.24) 1.1e-4 / 4.5e-4.  Gates: ~3-4x the
    # measurement (VERDICT round 1 item 1c asked for abs <= 1e-4, rel <= 4e-3).
    # tests/test_gpu_parity_at_size.py adds the gates normalised by the network's contribution.
    assert a < (3e-4 if cfg_name == "universe_original_16k" else 3e-5), (a, r)
    assert r < 2e-3, (a, r)


def test_full_size_properties():
    """BASELINE.json configs[1] shape (32 x 8 s) with a reduced step count: size-independent
    properties -- batch-shard invariance (rows are independent end to end),
    run-to-run determinism, shape / finiteness / peak limiter."""
    from open_universe_b200.config import builtin_config, instantiate
    torch.manual_seed(3)
    m = instantiate(builtin_config("universepp_16k").model, _recursive_=False)
    m.eval(no_ema=True)
    m = m.to(DEV)
    mix = det_audio((32, 128000), 99).to(DEV)

    def run(rows):
        rng = torch.Generator(device=DEV).manual_seed(1028282)
        # draw the GLOBAL noise and slice rows, as a batch-sharded rank does
        full = [torch.randn((32, 1, 128160), device=DEV, generator=rng) for _ in range(4)]
        from open_universe_b200.networks.universe import universe as U
        it = iter(full)
        old = U.randn
        U.randn = lambda x, sigma, rng=None: next(it)[rows] * sigma[:, None, None]
        try:
            return m.enhance(mix[rows], n_steps=4)
        finally:
            U.randn = old

    full = run(slice(0, 32))
    assert full.shape == (32, 128000) and torch.isfinite(full).all()
    assert float(full.abs().max()) <= 1.0 + 1e-6
    again = run(slice(0, 32))
    assert torch.equal(full, again)                      # deterministic kernels
    part = run(slice(8, 12))
    assert rel_rms(part.cpu(), full[8:12].cpu()) < 1e-6  # shard invariance


def test_load_model_roundtrip(tmp_path):
    """inference_utils.load_model on a reference-format checkpoint (state_dict + ema) + config.yaml."""
    import yaml
    from open_universe_b200 import inference_utils
    from open_universe_b200.config import CONFIG_DIR
    m = our_model("upp16k")
    raw = yaml.safe_load((CONFIG_DIR / "universepp_16k.yaml").read_text())
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(raw))
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    sd["loss_mpd.fake.weight"] = torch.zeros(3)          # training-only keys are ignored
    ema = {"decay": 0.999, "num_updates": 10, "collected_params": None,
           "shadow_params": [p.detach().cpu().clone() for p in m.model_parameters()]}
    torch.save({"state_dict": sd, "ema": ema}, tmp_path / "weights.ckpt")
    m2 = inference_utils.load_model(str(tmp_path / "weights.ckpt"), device=DEV)
    assert not m2.training and m2.fs == 16000
    mix = det_audio((1, 4000), 5).to(DEV)
    y1 = m.enhance(mix, n_steps=2, rng=torch.Generator(device=DEV).manual_seed(1))
    y2 = m2.enhance(mix, n_steps=2, rng=torch.Generator(device=DEV).manual_seed(1))
    assert torch.equal(y1, y2)


def test_cli_batched_directory(tmp_path):
    """bin/enhance.py end to end: checkpoint + config.yaml on disk, a folder of WAV files, batched by
    length; batch-size 1 reproduces a direct model.enhance() call with the CLI's seed."""
    import yaml
    from open_universe_b200.bin import enhance as cli
    from open_universe_b200.config import CONFIG_DIR
    m = our_model("upp16k")
    raw = yaml.safe_load((CONFIG_DIR / "universepp_16k.yaml").read_text())
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(raw))
    torch.save({"state_dict": {k: v.cpu() for k, v in m.state_dict().items()}}, tmp_path / "weights.ckpt")
    src, dst = tmp_path / "in", tmp_path / "out"
    (src / "sub").mkdir(parents=True)
    clips = {"a.wav": det_audio((1, 4000), 1), "sub/b.wav": det_audio((2, 4000), 2),
             "c.wav": det_audio((1, 2500), 3), "d.wav": det_audio((1, 4000), 4)}
    for name, x in clips.items():
        cli.save_audio(src / name, x, 16000)
    n = cli.main([str(src), str(dst), "--model", str(tmp_path / "weights.ckpt"), "--batch-size", "4",
                  "--n_steps", "2"])
    assert n == 4
    for name, x in clips.items():
        y, fs = cli.load_audio(dst / name)
        assert fs == 16000 and y.shape == x.shape and torch.isfinite(y).all()
    # one file per call == the reference's loop: same generator seed, same draw order
    one = tmp_path / "one.wav"
    cli.main([str(src / "a.wav"), str(one), "--model", str(tmp_path / "weights.ckpt"), "--n_steps", "2"])
    rng = torch.Generator(device=DEV).manual_seed(1028282)
    want = m.enhance(clips["a.wav"].to(DEV), n_steps=2, rng=rng).cpu()
    got, _ = cli.load_audio(one)
    assert torch.allclose(got, want, atol=1e-6)


def test_in_place_noise_draw_matches_reference_draw_order(monkeypatch):
    """enhance() draws the per-step noise straight into its persistent buffers (unit variance, sigma
    folded into the update coefficient) when ``randn`` is the stock function; a patched ``randn``
    takes the reference's literal path ``randn(x, sigma)``.  Same generator seed -> same noise, and the
    same result up to the re-association (beta * sigma) * n  vs  beta * (sigma * n)."""
    from open_universe_b200.config import builtin_config, instantiate
    from open_universe_b200.networks.universe import universe as U
    # the generator is consumed identically by randn(out=<slice of a buffer>) and randn(shape)
    buf = torch.empty(3, 2, 1, 4160, device=DEV)
    g1, g2 = (torch.Generator(device=DEV).manual_seed(5) for _ in range(2))
    for n in range(3):
        torch.randn(buf[n].shape, device=DEV, generator=g1, out=buf[n])
        assert torch.equal(buf[n], torch.randn(buf[n].shape, device=DEV, generator=g2))
    torch.manual_seed(4)
    m = instantiate(builtin_config("universepp_16k").model, _recursive_=False)   # reference init scheme
    m.eval(no_ema=True)
    m = m.to(DEV)
    mix = det_audio((2, 4000), 31).to(DEV)
    fast = m.enhance(mix, n_steps=5, rng=torch.Generator(device=DEV).manual_seed(7))
    stock = U.randn
    monkeypatch.setattr(U, "randn", lambda x, sigma, rng=None: stock(x, sigma, rng=rng))
    literal = m.enhance(mix, n_steps=5, rng=torch.Generator(device=DEV).manual_seed(7))
    err = rel_rms(fast.cpu(), literal.cpu())
    print("fast vs literal noise path: rel rms", err)
    assert err < 2e-3, err


@pytest.mark.parametrize("name,shape,level", [
    ("silence", (2, 1600), 0.0),          # zero variance: normalize_batch clamps the std at 1e-5 (norm.py:22-23)
    ("ten_samples", (1, 10), 0.05),       # shorter than one latent frame: T_pad = 160, ONE GRU step
    ("one_frame_exact", (3, 160), 0.05),  # exactly tot_ds: a full extra frame of padding (universe.py:221)
    ("odd_batch_ragged_tail", (5, 4001), 0.05),
])
def test_enhance_edge_cases_vs_oracle(name, shape, level, monkeypatch):
    """Degenerate inputs through the whole path against the CPU oracle (reference init scheme)."""
    from open_universe_b200.config import builtin_config, instantiate
    from oracle.universe_oracle import UniverseOracle
    torch.manual_seed(99)
    cfg = builtin_config("universepp_16k").model
    m = instantiate(cfg, _recursive_=False)
    m.eval(no_ema=True)
    o = UniverseOracle(cfg, m.state_dict())
    mix = det_audio(shape, 5, level=level) if level > 0 else torch.zeros(shape)
    t_pad = shape[-1] + (m.tot_ds - shape[-1] % m.tot_ds)
    n_steps = 3
    noise = det_noise(n_steps, (shape[0], 1, t_pad), 6)
    with torch.no_grad():
        want = o.enhance(mix, n_steps=n_steps, noise=noise)
    inject_noise(monkeypatch, noise)
    got = m.to(DEV).enhance(mix.to(DEV), n_steps=n_steps).cpu()
    assert got.shape == want.shape == mix.shape
    assert torch.isfinite(got).all()
    a = abs_rms(got, want)
    print(name, "abs rms err", a, "out rms", float(want.square().mean().sqrt()))
    assert a < 1e-3, a


def test_plan_replay_equals_per_launch_execution():
    """ScoreNetwork.forward through ONE ou_plan_run call == the same ops launched one by one from Python
    (Executor.run), bit for bit; also a partial replay (first, count) of the recorded list."""
    from open_universe_b200.engine import runtime as R
    m = our_model("upp16k")
    net = m.get_score_model()
    B, T = 2, 4800
    x = det_noise(1, (B, 1, T), 3)[0].to(DEV) * 0.3
    sigma = torch.tensor([0.3, 1.7], device=DEV)
    cond = [torch.randn(B, c, t, device=DEV, generator=torch.Generator(device=DEV).manual_seed(i))
            for i, (c, t) in enumerate([(512, 30), (256, 150), (128, 600), (64, 2400), (32, 4800)])]
    n0 = R.lib.launch_count()
    y_plan = net(x, sigma, cond)
    n_plan = R.lib.launch_count() - n0
    old = R.USE_PLAN
    R.USE_PLAN = False
    try:
        y_ops = net(x, sigma, cond)
    finally:
        R.USE_PLAN = old
    assert torch.equal(y_plan, y_ops)
    r = R.get_score_runner(net, B, T, x.device)
    assert r.__dict__.get("_plan") is not None and n_plan >= len(r.exe.prog.ops)
    (e0, ne), (g0, ng), (d0, nd) = r.phase_ranges()
    assert e0 == 0 and g0 == ne and d0 == ne + 1 and ne + ng + nd == len(r.exe.prog.ops)


def test_conditioner_plan_replay_equals_per_launch_execution():
    """ConditionerNetwork.forward (mel front-end included) through ONE ou_plan_run call == the same ops launched
    one by one from Python, bit for bit (ou_condition_forward)."""
    from open_universe_b200.engine import runtime as R
    m = our_model("upp16k")
    net = m.condition_model
    B, T = 2, 4800
    x = det_noise(1, (B, 1, T), 5)[0].to(DEV) * 0.1
    cond_p, y_p, h_p = net(x, x_wav=x, train=True)
    r = R.get_conditioner_runner(net, B, T, x.device, True)
    assert r.__dict__.get("_plan") is not None
    old = R.USE_PLAN
    R.USE_PLAN = False
    try:
        cond_o, y_o, h_o = net(x, x_wav=x, train=True)
    finally:
        R.USE_PLAN = old
    assert torch.equal(h_p, h_o) and torch.equal(y_p, y_o)
    assert all(torch.equal(a, b) for a, b in zip(cond_p, cond_o))
