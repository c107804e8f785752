"""Pin the CPU oracle (oracle/universe_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container)."""
import pytest
import torch

from cases import ENHANCE_CASES, NET_CASES, case_kwargs, noise_rows
from common import (abs_rms, det_audio, det_noise, golden_buffers, load_golden, make_oracle,
                    model_cfg, rel_rms, sub)

# fp32 re-association only (same library kernels, same op order): tight tolerance
TOL = 2e-5


@pytest.mark.parametrize("case", NET_CASES, ids=lambda c: c["name"])
def test_networks_match_reference(case):
    g = load_golden(case["name"])
    o = make_oracle(case["model"])
    B, T = case["B"], case["T"]
    x_wav = det_audio((B, 1, T), case["seed"], level=0.05)
    x_t = det_noise(1, (B, 1, T), case["seed"])[0] * 0.3
    sigma = torch.tensor((case["sigmas"] * B)[:B], dtype=torch.float32)
    with torch.no_grad():
        from oracle.universe_oracle import mel_spec
        mel = mel_spec(o.cfg["condition_model"], x_wav, o.sd, "condition_model.input_mel")
        cond, y_hat, h = o.condition(x_wav, x_wav)
        score = o.score(x_t, sigma, cond)
        net = o.net(x_t, sigma, cond)
    assert mel.shape == g["mel"].shape
    assert rel_rms(mel, g["mel"]) < TOL
    for name, t in [("y_hat", y_hat), ("h", h)] + [(f"cond{i}", c) for i, c in enumerate(cond)]:
        assert list(t.shape) == list(g[name + "_shape"]), name
        assert rel_rms(sub(t, g[name + "_stride"]), g[name]) < TOL, name
    assert rel_rms(net, g["net"]) < TOL
    assert rel_rms(score, g["score"]) < 5 * TOL  # (est - x)/sigma^2 amplifies at small sigma


@pytest.mark.parametrize("case", ENHANCE_CASES, ids=lambda c: c["name"])
def test_enhance_matches_reference(case):
    g = load_golden(case["name"])
    o = make_oracle(case["model"])
    shape = tuple(case["shape"])
    mix = det_audio(shape, case["seed"])
    noise = det_noise(case["n_steps"], (noise_rows(case), 1, int(g["t_pad"])), case["seed"])
    with torch.no_grad():
        y = o.enhance(mix, n_steps=case["n_steps"], noise=noise, **case_kwargs(case))
    assert y.shape == mix.shape == g["y"].shape
    assert rel_rms(y, g["y"]) < 10 * TOL, (rel_rms(y, g["y"]), abs_rms(y, g["y"]))


@pytest.mark.parametrize("model", ["upp16k", "orig16k", "upp24k"])
def test_constructor_buffers(model):
    """Binomial taps (blocks.py:62-68), Hann window and HTK mel filterbank (condition.py:75-81)
    rebuilt by the oracle equal the buffers the reference stores in its state_dict."""
    from oracle.universe_oracle import binomial_taps, mel_filterbank
    for k, v in golden_buffers(model).items():
        if k.endswith("low_pass_filter.weights"):
            assert torch.allclose(binomial_taps(v.numel()), v, atol=1e-6), k
        elif k.endswith("spectrogram.window"):
            assert torch.allclose(torch.hann_window(v.numel(), periodic=True), v, atol=1e-7), k
        elif k.endswith("mel_scale.fb"):
            assert torch.allclose(mel_filterbank(v.shape[0], v.shape[1]), v, atol=1e-6), k


def test_gru_explicit_matches_torch():
    from oracle.universe_oracle import gru_explicit
    torch.manual_seed(0)
    gru = torch.nn.GRU(16, 8, bidirectional=True, batch_first=True)
    x = torch.randn(3, 11, 16)
    with torch.no_grad():
        ref, _ = gru(x)
        f = gru_explicit(x, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0)
        b = gru_explicit(x, gru.weight_ih_l0_reverse, gru.weight_hh_l0_reverse,
                         gru.bias_ih_l0_reverse, gru.bias_hh_l0_reverse, reverse=True)
    assert torch.allclose(torch.cat([f, b], -1), ref, atol=1e-6)
