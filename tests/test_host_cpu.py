"""Host-side logic that needs no GPU: C-ABI library loads and exports every declared symbol,
config handling, drop-in surface (state_dict keys / EMA order / signatures), sampler tables,
checkpoint loading, and the loud failure of every compute entry point on CPU tensors."""
import argparse
import ctypes
import json
import math
import re
import typing
from pathlib import Path

import pytest
import torch

from common import GOLDEN, OUR_CONFIG, full_state_dict, golden_buffers, load_manifest, make_oracle
from open_universe_b200 import inference_utils
from open_universe_b200.config import CONFIG_DIR, builtin_config, instantiate, load_config
from open_universe_b200.engine import lib, runtime

ROOT = Path(__file__).resolve().parent.parent
_models = {}


def model(name):
    if name not in _models:
        _models[name] = instantiate(builtin_config(OUR_CONFIG[name]).model, _recursive_=False)
    return _models[name]


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "ou_b200.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|int64_t)\s+(ou_\w+)\s*\(", header, flags=re.M))
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    L = lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert L.ou_abi_version() == lib.OU_ABI_VERSION
    assert lib.launch_count() >= 0


def test_abi_rejects_bad_arguments_without_gpu():
    L = lib.load()
    assert L.ou_pack_blocked(None, None, 1, 8, 4, None) == -1
    assert "null" in lib.last_error()
    with pytest.raises(ValueError):
        lib.check(L.ou_gru_bidir(None, None, None, None, 1.0, None, 1, 1, 256, None))
    prm = lib.ConvParams()
    assert L.ou_conv1d(ctypes.byref(prm), None) == -1


@pytest.mark.parametrize("name", ["upp16k", "orig16k", "upp24k"])
def test_state_dict_matches_reference_layout(name):
    man = load_manifest(name)
    m = model(name)
    assert type(m).__name__ == man["class"]
    sd = m.state_dict()
    assert list(sd.keys()) == man["state_dict_order"]
    assert {k: list(v.shape) for k, v in sd.items()} == man["manifest"]
    by_id = {id(p): k for k, p in m.named_parameters()}
    assert [by_id[id(p)] for p in m.model_parameters()] == man["ema_param_order"]
    # constructor-built buffers equal the reference's (binomial taps, Hann window, mel fb)
    for k, v in golden_buffers(name).items():
        assert torch.allclose(sd[k], v, atol=1e-6), k
    m.load_state_dict(full_state_dict(name), strict=True)


def test_config_resolution_and_number_coercion():
    cfg = builtin_config("universe_original_16k")
    assert cfg.model.diffusion.sigma_min == 5e-4 and isinstance(cfg.model.diffusion.sigma_min, float)
    assert cfg.model.condition_model.rate_factors == [2, 4, 4, 5]
    assert cfg.model.condition_model.get("seq_model") == "gru"
    cfg24 = builtin_config("universepp_24k")
    assert cfg24.model.fs == 24000 and model("upp24k").tot_ds == 240


def test_enhance_signature_is_the_reference_api():
    m = model("upp16k")
    hints = typing.get_type_hints(m.enhance)
    assert list(hints) == ["n_steps", "epsilon", "target", "fake_score_snr", "rng", "use_aux_signal",
                           "keep_rms", "ensemble", "ensemble_stat", "warm_start", "return"]
    parser = inference_utils.add_enhance_arguments(m, argparse.ArgumentParser())
    args = parser.parse_args(["--n_steps", "16", "--keep_rms", "1"])
    assert args.n_steps == 16 and args.epsilon == 1.3 and args.keep_rms is True
    with pytest.raises(ValueError):
        inference_utils.add_enhance_arguments(object(), argparse.ArgumentParser())


def test_pad_unpad_rule():
    m = model("upp16k")
    for t in (1, 159, 160, 161, 32000, 128000):
        x = torch.zeros(1, 1, t)
        xp, pad = m.pad(x)
        assert pad == 160 - t % 160 and 1 <= pad <= 160
        assert xp.shape[-1] % 160 == 0 and xp.shape[-1] == t + pad
        assert m.unpad(xp, pad).shape[-1] == t


def test_sampler_tables_match_reference_formulas():
    """The affine (ca, cb, cc) form of the fused update equals universe.py:197-209,334-343."""
    torch.manual_seed(0)
    for name in ("upp16k", "orig16k"):
        m = model(name)
        o = make_oracle(name)
        n_steps, eps = 6, 1.3
        like = torch.zeros(1)
        sigma, net_sigma, in_scale, coef, (eta, beta) = m._sampler_tables(n_steps, eps, like)
        assert torch.allclose(sigma, o.sigmas(n_steps, like))
        x = torch.randn(2, 1, 50).double()
        net = torch.randn(2, 1, 50).double()
        z = torch.randn(2, 1, 50).double()
        for n in range(n_steps):
            s = sigma[n].double()
            if m.with_edm:
                sd = 10 ** (-26 / 20)
                est = sd**2 / (s**2 + sd**2) * x + s * sd / (s**2 + sd**2).sqrt() * net
                score = (est - x) / s**2
                assert math.isclose(in_scale[n].item(), 1 / math.sqrt(s**2 + sd**2), rel_tol=1e-6)
                assert math.isclose(net_sigma[n].item(), 0.25 * s.item(), rel_tol=1e-6)
            else:
                score = net
            if n < n_steps - 1:
                want = x + s**2 * eta * score + beta * z
            else:
                want = x + s**2 * score
            ca, cb, cc = coef[n].double()
            got = ca * x + cb * net + cc * z
            assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (name, n)


def test_load_model_checkpoint_formats(tmp_path):
    import yaml
    raw = yaml.safe_load((CONFIG_DIR / "universe_original_16k.yaml").read_text())
    run = tmp_path / "run"
    (run / ".hydra").mkdir(parents=True)
    (run / "checkpoints").mkdir()
    (run / ".hydra" / "config.yaml").write_text(yaml.safe_dump(raw))
    m = model("orig16k")
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    shadow = [torch.full_like(p, 0.5) for p in m.model_parameters()]
    ckpt = run / "checkpoints" / "last.ckpt"
    torch.save({"state_dict": sd, "ema": {"decay": 0.999, "num_updates": 3, "shadow_params": shadow,
                                          "collected_params": None}}, ckpt)
    m2, cfg = inference_utils.load_model(ckpt, device="cpu", return_config=True)
    assert cfg.model.fs == 16000 and not m2.training
    # eval() swapped the EMA shadow weights in (universe.py:849-855) ...
    assert all(bool((p == 0.5).all()) for p in m2.model_parameters())
    m2.train()   # ... and train() restores the raw ones
    for (k, p) in m2.named_parameters():
        assert torch.equal(p, sd[k]), k
    with pytest.raises(ValueError):
        inference_utils.load_model(tmp_path / "run" / "nothing_here" / "x.ckpt")


def test_no_cpu_fallback():
    m = model("upp16k")
    x = torch.zeros(1, 1, 320)
    with pytest.raises(runtime.NoCudaPathError):
        m.enhance(torch.zeros(1, 3200))
    with pytest.raises(runtime.NoCudaPathError):
        m.condition_model(x)
    with pytest.raises(runtime.NoCudaPathError):
        m.get_score_model()(x, torch.ones(1), [x])
    with pytest.raises(ValueError):
        m.enhance(torch.zeros(1, 1, 1, 320))
    with pytest.raises(runtime.NoCudaPathError):
        m.enhance(torch.zeros(1, 320), warm_start=2)
    with pytest.raises(ValueError):
        m.enhance(torch.zeros(1, 320), n_steps=4, warm_start=4)
    with pytest.raises(runtime.NoCudaPathError):
        m.enhance(torch.zeros(1, 320), target=torch.zeros(1, 320))
    with pytest.raises(NotImplementedError):      # universe.py:368
        m.enhance(torch.zeros(1, 320), ensemble=2, ensemble_stat="mode")


def test_signal_median_matches_oracle():
    from open_universe_b200.utils import signal_median
    from oracle.universe_oracle import signal_median as want
    g = torch.Generator().manual_seed(3)
    for e, b, s in [(3, 2, 50), (4, 3, 101), (5, 1, 7), (2, 2, 9)]:
        x = torch.randn(e, b, 1, s, generator=g)
        assert torch.equal(signal_median(x), want(x))


def test_film_channel_check():
    from open_universe_b200.networks.universe.blocks import film
    with pytest.raises(ValueError):
        film(torch.zeros(1, 4, 8), torch.zeros(1, 6))


def test_weights_version_tracks_inplace_updates():
    m = model("orig16k")
    v0 = runtime.weights_version(m.condition_model)
    with torch.no_grad():
        next(m.condition_model.parameters()).mul_(1.0)
    assert runtime.weights_version(m.condition_model) != v0


def test_weights_version_tracks_ema_swap():
    """eval() / train() / eval(no_ema=True) swap the live parameters with the EMA shadow
    (universe.py:841-865): the packed-weight cache key must change every time."""
    from open_universe_b200.engine import runtime
    torch.manual_seed(0)
    cfg = builtin_config("universepp_16k").model
    cfg["training"] = dict(cfg.get("training") or {}, ema_decay=0.999)
    m = instantiate(cfg, _recursive_=False)
    assert m.ema is not None
    with torch.no_grad():
        for s in m.ema.shadow_params:
            s.mul_(0.5)
    net = m.get_score_model()
    v_train = runtime.weights_version(net)
    m.eval()
    v_eval = runtime.weights_version(net)
    assert v_eval != v_train
    m.train()
    v_back = runtime.weights_version(net)
    assert v_back != v_eval
    m.eval(no_ema=True)
    assert runtime.weights_version(net) == v_back          # no swap, no change
    runtime.invalidate(net)
    assert runtime.weights_version(net) != v_back


def test_runner_cache_is_bounded_and_shares_packed_weights():
    """The per-module runner cache is a small LRU and keeps one packed-weight copy."""
    from open_universe_b200.engine import runtime

    class Dummy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(3))

    mod = Dummy()
    made = []

    def make(key):
        def f(shared):
            shared.setdefault("packed", object())
            made.append(key)
            return ("runner", key, shared["packed"])
        return f

    n = 2 * runtime.MAX_CACHED_SHAPES
    rs = [runtime._get_runner(mod, ("score", 1, t, "cpu"), make(t)) for t in range(n + 3)]
    cache = runtime._cache(mod)["runners"]
    assert len(cache) == n and ("score", 1, 0, "cpu") not in cache
    assert len({id(r[2]) for r in rs}) == 1                  # one shared packed-weight object
    runtime._get_runner(mod, ("score", 1, n + 2, "cpu"), make("again"))
    assert "again" not in made                                # cache hit
    with torch.no_grad():
        mod.w.add_(1.0)                                       # weights changed: everything is rebuilt
    runtime._get_runner(mod, ("score", 1, n + 2, "cpu"), make("rebuilt"))
    assert "rebuilt" in made and len(runtime._cache(mod)["runners"]) == 1


def test_plan_abi_records_ops_without_a_gpu():
    """Plan-level C ABI (include/ou_b200.h): create / add / size / destroy are host-only bookkeeping."""
    from ctypes import byref, c_void_p
    from open_universe_b200.engine import lib
    L = lib.load()
    plan = c_void_p()
    assert L.ou_plan_create(byref(plan)) == 0 and plan.value
    prm = lib.ConvParams()
    assert L.ou_plan_add_conv(plan, byref(prm), -1) == 0
    tp = lib.TrunkParams()
    assert L.ou_plan_add_trunk(plan, byref(tp), 0) == 0
    assert L.ou_plan_size(plan) == 2
    args = lib.StepArgs()
    assert L.ou_plan_run(plan, byref(args), 5, -1, None) == -1          # first op out of range -> OU_ERR_INVALID
    assert "out of range" in lib.last_error()
    assert L.ou_plan_run(plan, byref(args), 2, -1, None) == 0           # empty range: nothing launched
    assert L.ou_plan_destroy(plan) == 0


def test_gru_cta_count_and_plan_ops_for_the_conditioner():
    """Host-only parts of the round-2 ABI additions: ou_gru_ctas (what a host leaves free next to the overlapped
    recurrence) and the mel / GRU plan ops with their new arguments."""
    from ctypes import byref, c_void_p
    from open_universe_b200.engine import lib
    L = lib.load()
    # clusters of 8 CTAs per (direction, 8 clips); 4 clip slots when the batch fits; clusters of 4 only for H <= 256
    assert L.ou_gru_ctas(256, 32, 0) == 2 * 4 * 8 and L.ou_gru_ctas(256, 16, 4) == 2 * 2 * 4
    assert L.ou_gru_ctas(256, 3, 8) == 2 * 1 * 8 and L.ou_gru_ctas(384, 4, 4) == 2 * 1 * 8
    assert L.ou_gru_bidir_ex(None, None, None, None, 1.0, None, 1, 1, 256, 4, None) < 0     # null pointers are refused
    plan = c_void_p()
    assert L.ou_plan_create(byref(plan)) == 0
    one = c_void_p(16)      # any non-null pointer: recording does not dereference
    assert L.ou_plan_add_mel(plan, one, one, one, one, one, one, one, 2, 4800, 640, 160, 80, 240, 31) == 0
    assert L.ou_plan_add_gru(plan, one, one, one, None, 1.0, one, 2, 30, 256, 4) == 0
    assert L.ou_plan_add_mel(plan, None, one, one, one, one, one, one, 2, 4800, 640, 160, 80, 240, 31) < 0
    assert L.ou_plan_size(plan) == 2
    assert L.ou_plan_destroy(plan) == 0


def test_lowering_fuses_the_tails_and_hoists_the_up_prelus():
    """UNIVERSE++ 16 kHz: one evaluation is 31 launches -- enc.0.down, dec.4.up and the output conv ride on trunk
    launches, and the decoder's remaining up convs have no input PReLU of their own (engine/program.py)."""
    import torch
    from open_universe_b200.config import builtin_config, instantiate
    from open_universe_b200.engine import program as P
    torch.manual_seed(0)
    m = instantiate(builtin_config("universepp_16k").model, _recursive_=False)
    prog = P.lower_score_network(m.get_score_model(), 2, 16000)
    trunks = {op.name: op for op in prog.ops if isinstance(op, P.TrunkOp)}
    assert len(prog.ops) == 31 and set(trunks) == {"enc.0.trunk", "enc.1.trunk", "dec.3.trunk", "dec.4.trunk"}
    assert trunks["enc.0.trunk"].tail_dn is not None and trunks["enc.0.trunk"].tail_dn.fc.taps == 3
    assert trunks["dec.3.trunk"].tail is not None and trunks["dec.4.trunk"].tail_out is not None
    ups = [op for op in prog.ops if isinstance(op, P.ConvOp) and op.fc.up > 1]
    assert [u.name for u in ups] == ["dec.1.up", "dec.2.up", "dec.3.up"] and all(u.fc.prelu_in is None for u in ups)
    by_name = {op.name: op for op in P.flat_ops(prog.ops)}
    assert all(by_name[n].prelu_out is not None for n in ("dec.0.conv3", "dec.1.conv3", "dec.2.conv3"))
    # UNIVERSE (original): plain k = s rate convs take the same tails through their 1-tap form
    m = instantiate(builtin_config("universe_original_16k").model, _recursive_=False)
    prog = P.lower_score_network(m.get_score_model(), 2, 16000)
    trunks = {op.name: op for op in prog.ops if isinstance(op, P.TrunkOp)}
    assert trunks["enc.0.trunk"].tail_dn.fc.taps == 1 and trunks["dec.3.trunk"].tail.fc.taps == 1
