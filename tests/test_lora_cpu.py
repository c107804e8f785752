"""LoRA surface on CPU: state_dict layout against the manifest dumped from the real reference's
UniverseLoRA, exactness of the load-time merge + lowering against the reference's golden output, and the
oracle's partial_diffusion port against the reference golden (pins the oracle for the GPU test)."""
import json

import torch

from cases import LORA_CASES, LORA_RANK, LORA_SEED
from common import GOLDEN, OUR_CONFIG, det_audio, det_noise, full_state_dict, load_golden, rel_rms
from detweights import det_lora_factors


def build_lora_model(case, device="cpu"):
    from open_universe_b200.config import builtin_config, instantiate
    from open_universe_b200.networks.universe import UniverseLoRA
    m = instantiate(builtin_config(OUR_CONFIG[case["model"]]).model, _recursive_=False)
    m.load_state_dict(full_state_dict(case["model"]), strict=True)
    if m.ema is not None:     # the wrapper copies the EMA shadow in (lora.py:143-146): make it the loaded weights
        m.ema.shadow_params = [p.clone().detach() for p in m.model_parameters()]
    L = UniverseLoRA(m, m.fs, diffusion={"n_steps": case["n_steps"], "epsilon": 1.3}, lora_rank=LORA_RANK,
                     use_partial_diffusion=case["partial"])
    shapes = {k: list(v.shape) for k, v in L.state_dict().items() if ".lora_" in k}
    missing, unexpected = L.load_state_dict(det_lora_factors(shapes, LORA_SEED), strict=False)
    assert not unexpected
    L.eval()
    return L.to(device)


def merged_oracle(L):
    """CPU oracle of the adapted model: adapters replaced by plain layers holding the merged weights."""
    import copy
    from open_universe_b200 import lora
    from open_universe_b200.config import builtin_config
    from oracle.universe_oracle import UniverseOracle
    plain = copy.deepcopy(L.model).cpu()
    lora.remove(plain)
    return UniverseOracle(builtin_config("universepp_16k").model, plain.state_dict())


def test_lora_state_dict_matches_reference_layout():
    man = json.loads((GOLDEN / "upp16k_lora_manifest.json").read_text())
    L = build_lora_model(LORA_CASES[0])
    sd = L.state_dict()
    assert list(sd) == man["state_dict_order"]
    assert {k: list(v.shape) for k, v in sd.items()} == man["manifest"]
    trainable = {k for k, p in L.named_parameters() if p.requires_grad}
    assert all(".lora_" in k or "bias" in k for k in trainable) and any(".lora_" in k for k in trainable)


def test_lora_merge_and_lowering_exact_vs_reference_golden():
    """The adapted model lowered with the merged weights and emulated in exact fp32 reproduces the
    reference's UniverseLoRA.forward (full sampler) to fp32 round-off."""
    from emul_enhance import emulated_enhance
    case = LORA_CASES[0]
    g = load_golden(case["name"])
    L = build_lora_model(case)
    o = merged_oracle(L)
    mix = det_audio(tuple(case["shape"]), case["seed"])
    t = case["shape"][-1]
    t_pad = t + (L.model.tot_ds - t % L.model.tot_ds)
    noise = det_noise(case["n_steps"], (case["shape"][0], 1, t_pad), case["seed"])
    with torch.no_grad():
        want = o.enhance(mix, n_steps=case["n_steps"], noise=noise)
        got = emulated_enhance(L.model, o, mix, case["n_steps"], noise, False)
    assert rel_rms(want, g["y"]) < 2e-5          # oracle on merged weights == reference with adapters
    assert rel_rms(got, g["y"]) < 5e-5           # our fold (LoRA merge included) + lowering


def test_oracle_partial_diffusion_vs_reference_golden():
    case = LORA_CASES[1]
    g = load_golden(case["name"])
    L = build_lora_model(case)
    o = merged_oracle(L)
    B, T = case["shape"]
    mix = det_audio((B, T), case["seed"])
    noise = det_noise(case["n_steps"], (B, 1, T), case["seed"])
    torch.manual_seed(case["seed"])
    t_final = torch.zeros(B).uniform_(0, 1)       # the reference's draw (lora.py:246), same generator state
    with torch.no_grad():
        y = o.partial_diffusion(mix[:, None, :], case["n_steps"], 1.3, t_final, noise)[:, 0, :]
    assert y.shape == g["y"].shape
    assert rel_rms(y, g["y"]) < 2e-5
