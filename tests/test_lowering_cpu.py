"""Host-side lowering (weight folding, epilogue wiring, buffer lengths) checked on CPU: the op
list produced by ``engine.program`` is executed by the PyTorch emulator (oracle/emulator.py,
exact fp32 mode) and compared with the oracle.  No GPU, no CUDA library involved."""
import math

import pytest
import torch

from cases import NET_CASES
from common import det_audio, det_noise, full_state_dict, make_oracle, model_cfg, rel_rms
from oracle import emulator as E
from open_universe_b200.config import instantiate
from open_universe_b200.engine import program as P

_models = {}


def our_model(name):
    if name not in _models:
        m = instantiate(model_cfg_full(name), _recursive_=False)
        missing, unexpected = m.load_state_dict(full_state_dict(name), strict=True)
        m.eval(no_ema=True)
        _models[name] = m
    return _models[name]


def model_cfg_full(name):
    from common import OUR_CONFIG
    from open_universe_b200.config import builtin_config
    return builtin_config(OUR_CONFIG[name]).model


def run_case(case, quant):
    o = make_oracle(case["model"])
    m = our_model(case["model"])
    B, T = case["B"], case["T"]
    x_wav = det_audio((B, 1, T), case["seed"], level=0.05)
    x_t = det_noise(1, (B, 1, T), case["seed"])[0] * 0.3
    sigma = torch.tensor((case["sigmas"] * B)[:B], dtype=torch.float32)
    with torch.no_grad():
        cond_ref, y_ref, h_ref = o.condition(x_wav, x_wav)
        net_ref = o.net(x_t, sigma, cond_ref)
        # conditioner
        cp = P.lower_conditioner(m.condition_model, B, T)
        bufs, _, _ = E.run_program(cp, {"x": x_wav, "x_wav": x_wav}, quant=quant)
        cond = [bufs[cp.outputs[f"cond{i}"]] for i in range(len(cond_ref))]
        errs = {"h": rel_rms(bufs[cp.outputs["h"]], h_ref),
                "y_hat": rel_rms(bufs[cp.outputs["y_hat"]], y_ref[..., : cp.meta["t_final"]])}
        for i, (a, b) in enumerate(zip(cond, cond_ref)):
            assert a.shape == b.shape
            errs[f"cond{i}"] = rel_rms(a, b)
        # score network on the ORACLE's conditioning (isolates the two lowerings)
        net = m.get_score_model()
        sp = P.lower_score_network(net, B, T)
        pp = P.lower_cond_projection(net, B, sp.meta["lengths"])
        pb, _, _ = E.run_program(pp, {f"cond{i}": E._q(c, quant) for i, c in enumerate(cond_ref)},
                                 quant=quant)
        g = o_sigma_embedding(o, sigma)
        film = E.film_table(sp, g)
        inputs = {"x": x_t}
        inputs.update({k: v for k, v in pb.items() if k.startswith("sc")})
        _, net_out, _ = E.run_program(sp, inputs, film=film, quant=quant)
        errs["net"] = rel_rms(net_out, net_ref)
    return errs


def o_sigma_embedding(o, sigma):
    from oracle.universe_oracle import sigma_embedding
    return sigma_embedding(o.cfg["score_model"], o.sd, o.score_prefix + ".sigma_block",
                           torch.log10(sigma))


@pytest.mark.parametrize("case", NET_CASES, ids=lambda c: c["name"])
def test_lowering_exact(case):
    errs = run_case(case, quant=False)
    assert max(errs.values()) < 2e-5, errs


@pytest.mark.parametrize("case", [NET_CASES[0], NET_CASES[2]], ids=lambda c: c["name"])
def test_lowering_storage_policy(case):
    """fp16 storage / operands, fp32 accumulate (the library default): the error of a single
    evaluation on the unit-gain stress weights stays below 5e-3 relative (bf16: ~2e-2)."""
    from oracle import emulator as E
    assert E.QDTYPE == torch.float16
    errs = run_case(case, quant=True)
    print(errs)
    assert max(errs.values()) < 5e-3, errs
