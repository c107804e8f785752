"""Generate the committed golden vectors by running the UNMODIFIED reference on CPU.

Run by hand in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/<model>_manifest.json`` (state_dict key -> shape, buffers),
``tests/golden/<model>_buffers.npz`` (constructor-built buffers) and
``tests/golden/<case>.npz`` (outputs of the reference for each case in ``cases.py``).
Inputs, weights and diffusion noise are NOT stored: they are regenerated from seeds by
``detweights.py`` on every box.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ref_shim  # noqa: E402
from cases import (ENHANCE_CASES, LORA_CASES, LORA_RANK, LORA_SEED, MODELS, NET_CASES, WEIGHT_SEED,  # noqa: E402
                   case_kwargs, noise_rows)
from detweights import (det_audio, det_lora_factors, det_noise, det_state_dict,  # noqa: E402
                        is_constructor_buffer, subsample)

torch.set_num_threads(8)
_models = {}


def _named_model_parameters(model):
    """Names in the order of model.model_parameters()."""
    by_id = {id(p): k for k, p in model.named_parameters()}
    return [(by_id[id(p)], p) for p in model.model_parameters()]


def get_model(name):
    if name in _models:
        return _models[name]
    model, cfg = ref_shim.build_reference_model(MODELS[name])
    sd = model.state_dict()
    manifest = {k: list(v.shape) for k, v in sd.items() if not k.startswith("loss_")}
    param_names = {k for k, _ in model.named_parameters()}
    meta = {
        "reference_yaml": MODELS[name],
        "class": type(model).__name__,
        "manifest": manifest,
        "buffers": sorted(k for k in manifest if k not in param_names),
        # order matters: torch_ema stores its shadow params as a positional list in the order
        # of Universe.model_parameters() (universe.py:130-133, universe_gan.py:136-143)
        "ema_param_order": [k for k, _ in _named_model_parameters(model)],
        "state_dict_order": [k for k in sd if not k.startswith("loss_")],
        "n_loss_keys": sum(k.startswith("loss_") for k in sd),
    }
    (HERE / f"{name}_manifest.json").write_text(json.dumps(meta, indent=0, sort_keys=True))
    np.savez(HERE / f"{name}_buffers.npz",
             **{k: sd[k].numpy() for k in manifest if is_constructor_buffer(k)})
    new = det_state_dict(manifest, WEIGHT_SEED)
    missing, unexpected = model.load_state_dict(new, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("loss_") or is_constructor_buffer(k) for k in missing), missing
    # the EMA shadow is what eval() copies into the live params (universe.py:849-855):
    # make it equal to the freshly loaded weights
    if model.ema is not None:
        model.ema.shadow_params = [p.clone().detach() for p in model.model_parameters()]
    model.eval()
    _models[name] = (model, cfg)
    return _models[name]


def run_enhance_case(case):
    model, _ = get_model(case["model"])
    shape = tuple(case["shape"])
    mix = det_audio(shape, case["seed"])
    b = noise_rows(case)
    t = shape[-1]
    t_pad = t + (model.tot_ds - t % model.tot_ds)
    noise = det_noise(case["n_steps"], (b, 1, t_pad), case["seed"])
    ref_shim.set_injected_noise(noise)
    with torch.no_grad():
        y = model.enhance(mix, n_steps=case["n_steps"], **case_kwargs(case))
    assert y.shape == mix.shape
    np.savez(HERE / f"{case['name']}.npz", y=y.numpy(), t_pad=np.int64(t_pad))
    print(case["name"], "rms", float(y.square().mean().sqrt()))


def run_net_case(case):
    model, _ = get_model(case["model"])
    B, T = case["B"], case["T"]
    x_wav = det_audio((B, 1, T), case["seed"], level=0.05)
    x_t = det_noise(1, (B, 1, T), case["seed"])[0] * 0.3
    sigma = torch.tensor((case["sigmas"] * B)[:B], dtype=torch.float32)
    out = {}
    with torch.no_grad():
        mel = model.condition_model.input_mel.compute_mel_spec(x_wav)
        cond, y_hat, h = model.condition_model(x_wav, x_wav=x_wav, train=True)
        score = model.score_model(x_t, sigma, cond)
        net = model.get_score_model()(x_t, sigma, cond)
    out["mel"] = mel.numpy()
    out["score"] = score.numpy()
    out["net"] = net.numpy()
    for name, t in [("y_hat", y_hat), ("h", h)] + [(f"cond{i}", c) for i, c in enumerate(cond)]:
        sub, stride = subsample(t)
        out[name] = sub.numpy()
        out[name + "_stride"] = np.int64(stride)
        out[name + "_shape"] = np.array(t.shape, dtype=np.int64)
        out[name + "_rms"] = np.float64(t.double().square().mean().sqrt())
    np.savez(HERE / f"{case['name']}.npz", **out)
    print(case["name"], "score rms", float(score.square().mean().sqrt()),
          "cond shapes", [tuple(c.shape) for c in cond])


def run_lora_case(case):
    """Reference UniverseLoRA (lora.py) around a freshly built base model with the deterministic weights."""
    import importlib
    get_model(case["model"])                     # writes the manifests
    _models.pop(case["model"])                   # UniverseLoRA rewrites its base model: use a private copy
    base, _ = get_model(case["model"])
    _models.pop(case["model"])
    ref_shim.install()
    lora_mod = importlib.import_module("open_universe.networks.universe.lora")
    n_steps = case["n_steps"]
    m = lora_mod.UniverseLoRA(
        model=base, fs=base.fs, losses=ref_shim.to_attr({}), training=ref_shim.to_attr({"ema_decay": 0.0}),
        validation=ref_shim.to_attr({"enh_losses": {}}), optimizer=None, scheduler=None, grad_clipper=None,
        diffusion=ref_shim.to_attr({"n_steps": n_steps, "epsilon": 1.3}), lora_rank=LORA_RANK,
        use_partial_diffusion=case["partial"])
    m.eval()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    lora_keys = {k: v for k, v in shapes.items() if ".lora_" in k}
    (HERE / f"{case['model']}_lora_manifest.json").write_text(
        json.dumps({"manifest": shapes, "state_dict_order": list(shapes)}, indent=0, sort_keys=True))
    missing, unexpected = m.load_state_dict(det_lora_factors(lora_keys, LORA_SEED), strict=False)
    assert not unexpected
    shape = tuple(case["shape"])
    mix = det_audio(shape, case["seed"])
    t = shape[-1]
    t_len = t if case["partial"] else t + (base.tot_ds - t % base.tot_ds)
    noise = det_noise(n_steps, (shape[0], 1, t_len), case["seed"])
    ref_shim.set_injected_noise(noise)
    lora_mod.randn = importlib.import_module("open_universe.networks.universe.universe").randn
    torch.manual_seed(case["seed"])          # t_final = zeros(B).uniform_(0, 1) comes from the global generator
    with torch.no_grad():
        y = m(mix, n_steps=n_steps)
    y = y.detach()       # partial_diffusion re-enables grad for its last steps (lora.py:272, 290)
    np.savez(HERE / f"{case['name']}.npz", y=y.numpy())
    print(case["name"], "rms", float(y.square().mean().sqrt()), "shape", tuple(y.shape))


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for c in LORA_CASES:
        if not only or c["name"] in only:
            run_lora_case(c)
    if only and only <= {c["name"] for c in LORA_CASES}:
        sys.exit(0)
    for c in NET_CASES:
        if not only or c["name"] in only:
            run_net_case(c)
    for c in ENHANCE_CASES:
        if not only or c["name"] in only:
            run_enhance_case(c)
