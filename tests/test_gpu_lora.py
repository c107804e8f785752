"""UniverseLoRA.forward / partial_diffusion on the CUDA path (forward only) against the golden vectors
of the real reference (deterministic stress weights: loose, wiring-level) and against the CPU oracle on
the merged weights with the reference's init scheme (tight)."""
import pytest
import torch

from cases import LORA_CASES, LORA_RANK, LORA_SEED
from common import abs_rms, det_audio, det_noise, load_golden, rel_rms
from detweights import det_lora_factors
from test_lora_cpu import build_lora_model, merged_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def inject_noise(monkeypatch, noise):
    from open_universe_b200.networks.universe import universe as U
    it = iter(noise)

    def randn(x, sigma, rng=None):
        n = next(it).to(x)
        assert n.shape == x.shape
        return n * sigma[:, None, None]

    monkeypatch.setattr(U, "randn", randn)


@pytest.mark.parametrize("case", LORA_CASES, ids=lambda c: c["name"])
def test_lora_forward_vs_reference_golden(case, monkeypatch):
    g = load_golden(case["name"])
    L = build_lora_model(case, DEV)
    B, T = case["shape"]
    mix = det_audio((B, T), case["seed"])
    t_len = T if case["partial"] else T + (L.model.tot_ds - T % L.model.tot_ds)
    noise = det_noise(case["n_steps"], (B, 1, t_len), case["seed"])
    inject_noise(monkeypatch, noise)
    if case["partial"]:
        torch.manual_seed(case["seed"])
        t_final = torch.zeros(B).uniform_(0, 1).to(DEV)
        y = L.partial_diffusion(mix[:, None, :].to(DEV), t_final=t_final)[:, 0, :].cpu()
    else:
        y = L(mix.to(DEV), n_steps=case["n_steps"]).cpu()
    assert y.shape == g["y"].shape and torch.isfinite(y).all()
    err = rel_rms(y, g["y"])
    print(case["name"], "rel", err)
    assert err < 0.12, err          # stress weights (tests/test_gpu_networks.py docstring)


@pytest.mark.parametrize("partial", [False, True], ids=["full", "partial"])
def test_lora_forward_vs_oracle_init_weights(partial, monkeypatch, parity_log):
    from open_universe_b200.config import builtin_config, instantiate
    from open_universe_b200.networks.universe import UniverseLoRA
    torch.manual_seed(4321)
    m = instantiate(builtin_config("universepp_16k").model, _recursive_=False)
    n_steps = 6
    L = UniverseLoRA(m, m.fs, diffusion={"n_steps": n_steps, "epsilon": 1.3}, lora_rank=LORA_RANK,
                     use_partial_diffusion=partial)
    shapes = {k: list(v.shape) for k, v in L.state_dict().items() if ".lora_" in k}
    L.load_state_dict(det_lora_factors(shapes, LORA_SEED), strict=False)
    L.eval()
    o = merged_oracle(L)
    B, T = 2, 16000
    mix = det_audio((B, T), 61)
    t_len = T if partial else T + (m.tot_ds - T % m.tot_ds)
    noise = det_noise(n_steps, (B, 1, t_len), 61)
    t_final = torch.tensor([0.15, 0.6])
    with torch.no_grad():
        if partial:
            want = o.partial_diffusion(mix[:, None, :], n_steps, 1.3, t_final, noise)[:, 0, :]
        else:
            want = o.enhance(mix, n_steps=n_steps, noise=noise)
    inject_noise(monkeypatch, noise)
    L = L.to(DEV)
    if partial:
        got = L.partial_diffusion(mix[:, None, :].to(DEV), t_final=t_final.to(DEV))[:, 0, :].cpu()
    else:
        got = L(mix.to(DEV), n_steps=n_steps).cpu()
    a, r = abs_rms(got, want), rel_rms(got, want)
    parity_log[f"lora_{'partial' if partial else 'full'}_universepp_16k_{B}x{T}_{n_steps}"] = {
        "abs_rms_err": a, "rel_rms_err": r, "out_rms": float(want.square().mean().sqrt())}
    print("lora", "partial" if partial else "full", "abs", a, "rel", r)
    assert got.shape == want.shape
    assert a < 1e-4 and r < 4e-3, (a, r)


def test_lora_factor_update_repacks_weights():
    """In-place updates of the LoRA factors (an optimiser step) must reach the packed device weights."""
    case = LORA_CASES[0]
    L = build_lora_model(case, DEV)
    mix = det_audio((1, 2400), 3).to(DEV)
    y0 = L(mix, n_steps=2, rng=torch.Generator(device=DEV).manual_seed(1))
    with torch.no_grad():
        for k, p in L.named_parameters():
            if k.endswith("lora_weight_a"):
                p.mul_(1.5)
    y1 = L(mix, n_steps=2, rng=torch.Generator(device=DEV).manual_seed(1))
    assert not torch.equal(y0, y1)
