"""Value parity of the CUDA path against the CPU oracle AT the sizes of BASELINE.json's configs and
with gates that a wrong network cannot pass.

Weights come from the reference's own initialisation scheme (our constructors reproduce the
reference's ``state_dict`` layout and init, seeded).  With those weights the score network's
contribution to the output of ``enhance()`` is small (replacing it by zeros moves the output by
~11 % relative), so next to the north_star gate (absolute RMS) every end-to-end test here also gates

  * the error NORMALISED BY THE NETWORK'S CONTRIBUTION  ||got - want|| / ||want - want(zero net)||,
  * a weight set where the network matters as much as the signal ("out16": the score network's
    output conv scaled x16 so that net(x) has about unit RMS, the regime of a trained EDM network),
  * the relative error of ONE raw network evaluation at three noise levels.

Every measured error is recorded through the ``parity_log`` fixture and written to
``gpurun_out/parity_r2.json`` at the end of the session (copied to ``profiles/`` by hand).
"""
import pytest
import torch

from common import abs_rms, det_audio, det_noise, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build(cfg_name, seed=1234, out_gain=1.0):
    from open_universe_b200.config import builtin_config, instantiate
    from oracle.universe_oracle import UniverseOracle
    torch.manual_seed(seed)
    cfg = builtin_config(cfg_name).model
    m = instantiate(cfg, _recursive_=False)
    m.eval(no_ema=True)
    if out_gain != 1.0:
        with torch.no_grad():
            m.get_score_model().output_conv.conv.weight_g.mul_(out_gain)
    return m, UniverseOracle(cfg, m.state_dict())


def inject_noise(monkeypatch, noise):
    from open_universe_b200.networks.universe import universe as U
    it = iter(noise)

    def randn(x, sigma, rng=None):
        n = next(it).to(x)
        assert n.shape == x.shape
        return n * sigma[:, None, None]

    monkeypatch.setattr(U, "randn", randn)


def oracle_enhance(o, mix, n_steps, noise, zero_net=False):
    import oracle.universe_oracle as UO
    real = UO.score_network
    if zero_net:
        UO.score_network = lambda cfg, sd, p, x, sigma, cond: torch.zeros_like(x)
    try:
        with torch.no_grad():
            return o.enhance(mix, n_steps=n_steps, noise=noise)
    finally:
        UO.score_network = real


def run_case(monkeypatch, cfg_name, shape, n_steps, out_gain=1.0, with_zero=False, seed=77):
    m, o = build(cfg_name, out_gain=out_gain)
    mix = det_audio(shape, seed)
    t_pad = shape[-1] + (m.tot_ds - shape[-1] % m.tot_ds)
    noise = det_noise(n_steps, (shape[0], 1, t_pad), seed)
    want = oracle_enhance(o, mix, n_steps, noise)
    inject_noise(monkeypatch, noise)
    got = m.to(DEV).enhance(mix.to(DEV), n_steps=n_steps).cpu()
    assert got.shape == want.shape and torch.isfinite(got).all()
    res = {"config": cfg_name, "shape": list(shape), "n_steps": n_steps, "out_gain": out_gain,
           "abs_rms_err": abs_rms(got, want), "rel_rms_err": rel_rms(got, want),
           "out_rms": float(want.square().mean().sqrt())}
    if with_zero:
        zero = oracle_enhance(o, mix, n_steps, noise, zero_net=True)
        contrib = abs_rms(want, zero)
        res["net_contribution_rms"] = contrib
        res["err_over_contribution"] = res["abs_rms_err"] / contrib
    return res


def abs_gate(cfg_name):
    return 3e-4 if cfg_name == "universe_original_16k" else 3e-5


# BASELINE.json configs[1..3] at their clip length and step count (batch reduced to what the CPU oracle
# finishes in under a minute on the GPU box's host; rows are independent, see test_full_size_properties)
AT_SIZE = [
    ("cfg2_upp16k_2x8s_64steps", "universepp_16k", (2, 128000), 64),
    ("cfg3_orig16k_2x4s_32steps", "universe_original_16k", (2, 64000), 32),
    ("cfg4_upp24k_1x10s_64steps", "universepp_24k", (1, 240000), 64),
    ("cfg4_upp24k_4x10s_4steps", "universepp_24k", (4, 240000), 4),     # cfg-4's per-GPU batch (B/R = 4)
]


@pytest.mark.timeout(900, method="thread")
@pytest.mark.parametrize("name,cfg_name,shape,n_steps", AT_SIZE, ids=[c[0] for c in AT_SIZE])
def test_enhance_at_baseline_size(name, cfg_name, shape, n_steps, monkeypatch, parity_log):
    r = run_case(monkeypatch, cfg_name, shape, n_steps)
    parity_log[name] = r
    print(name, r)
    # north_star: <= 1e-3 RMS.  Measured (profiles/parity_r2.json): UNIVERSE++ 16k / 24k 3.6e-6 .. 4.9e-6
    # absolute, 8e-5 .. 2.5e-4 relative; UNIVERSE original (no EDM wrapper, output RMS 0.17) 5.7e-5 / 3.4e-4.
    # Gates at ~4x the measurement.
    assert r["abs_rms_err"] < abs_gate(cfg_name), r
    assert r["rel_rms_err"] < 2e-3, r


@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("cfg_name,shape,n_steps", [
    ("universepp_16k", (1, 16000), 8),
    ("universepp_16k", (2, 32000), 16),
    ("universe_original_16k", (1, 16000), 8),
    ("universepp_24k", (1, 24000), 8),
])
def test_error_normalised_by_network_contribution(cfg_name, shape, n_steps, monkeypatch, parity_log):
    """||got - want|| / ||want - want(score network := 0)||: a network that is x % wrong fails at x %."""
    r = run_case(monkeypatch, cfg_name, shape, n_steps, with_zero=True)
    parity_log[f"contribution_{cfg_name}_{shape[0]}x{shape[1]}_{n_steps}"] = r
    print(r)
    assert r["net_contribution_rms"] > 1e-3          # the denominator is not degenerate
    assert r["err_over_contribution"] < 5e-3, r      # VERDICT asks <= 2e-2; measured 3e-4 .. 1.2e-3
    assert r["abs_rms_err"] < abs_gate(cfg_name), r


@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("cfg_name,shape,n_steps", [
    ("universepp_16k", (1, 16000), 8),
    ("universepp_16k", (2, 64000), 32),
    ("universepp_24k", (1, 24000), 8),
])
def test_enhance_with_network_dominated_weights(cfg_name, shape, n_steps, monkeypatch, parity_log):
    """'out16' weights: the score network's output conv x16 -> net(x) ~ unit RMS, its contribution to the
    output is as large as the output itself; every internal activation is unchanged."""
    r = run_case(monkeypatch, cfg_name, shape, n_steps, out_gain=16.0, with_zero=True)
    parity_log[f"out16_{cfg_name}_{shape[0]}x{shape[1]}_{n_steps}"] = r
    print(r)
    assert r["net_contribution_rms"] > 0.3 * r["out_rms"], r
    # measured: 16k 5e-5 .. 6e-5 absolute (output RMS 0.05), 24k 2.4e-4 (output RMS 0.67); 8e-4 .. 1.2e-3
    # of the network's contribution
    assert r["abs_rms_err"] < 1e-3, r                # north_star
    assert r["rel_rms_err"] < 4e-3, r
    assert r["err_over_contribution"] < 5e-3, r


@pytest.mark.timeout(600, method="thread")
@pytest.mark.parametrize("cfg_name,seconds", [("universepp_16k", 8.0), ("universe_original_16k", 4.0),
                                              ("universepp_24k", 10.0)])
def test_per_evaluation_network_error(cfg_name, seconds, parity_log):
    """ONE raw ScoreNetwork evaluation at a BASELINE clip length against the oracle, at a high, a middle
    and the lowest noise level of the schedule, on the oracle's conditioning (isolates the score
    network from conditioner error) and on our own (the full chain)."""
    m, o = build(cfg_name)
    T = int(m.fs * seconds)
    B = 2
    mix = det_audio((B, 1, T), 5)
    mixp, _ = m.pad(mix)
    mixn = o.normalize(mixp)
    z = det_noise(1, tuple(mixn.shape), 6)[0]
    with torch.no_grad():
        cond_ref, _, _ = o.condition(mixn, mixn)
    m = m.to(DEV)
    net = m.get_score_model()
    cond_own, _, _ = m.condition_model(mixn.to(DEV), x_wav=mixn.to(DEV), train=True)
    worst = 0.0
    for sig in (2.0, 0.1, 0.005):
        sigma = torch.full((B,), sig)
        x = mixn + sig * z
        with torch.no_grad():
            want = o.net(x, sigma, cond_ref)
        got = net(x.to(DEV), sigma.to(DEV), [c.to(DEV) for c in cond_ref]).cpu()
        got_own = net(x.to(DEV), sigma.to(DEV), cond_own).cpu()
        e, e_own = rel_rms(got, want), rel_rms(got_own, want)
        parity_log[f"per_eval_{cfg_name}_sigma{sig}"] = {
            "config": cfg_name, "seconds": seconds, "sigma": sig, "rel_rms_err_oracle_cond": e,
            "rel_rms_err_own_cond": e_own, "net_rms": float(want.square().mean().sqrt())}
        print(cfg_name, "sigma", sig, "per-evaluation rel rms", e, "with own conditioning", e_own)
        worst = max(worst, e, e_own)
    # bf16 storage measured 2.6e-2; the bar set for round 2 is <= 1e-2 (about 2x the
    # reference's own TF32 arithmetic); fp16 storage measures 1e-3 .. 6e-3
    assert worst < 1e-2, worst
