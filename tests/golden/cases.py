"""Golden-vector case list shared by ``make_golden.py`` (generator) and the parity tests."""

# our model name -> reference ``config/model/<file>.yaml``
MODELS = {
    "upp16k": "default",             # UNIVERSE++ 16 kHz  (config/model/default.yaml)
    "orig16k": "universe_original",  # UNIVERSE 16 kHz    (config/model/universe_original.yaml)
    "upp24k": "universepp_24k",      # UNIVERSE++ 24 kHz  (config/model/universepp_24k.yaml)
}

WEIGHT_SEED = 7

# enhance() end-to-end with injected diffusion noise (universe.py:231-375)
ENHANCE_CASES = [
    dict(name="upp16k_short", model="upp16k", shape=(1, 8000), n_steps=4, seed=11, kwargs={}),
    # BASELINE.json configs[0]: 1 clip x 2 s x 8 steps
    dict(name="upp16k_cfg1", model="upp16k", shape=(1, 32000), n_steps=8, seed=12, kwargs={}),
    # length already a multiple of tot_ds=160 -> a FULL extra 160 of padding (universe.py:221)
    dict(name="upp16k_b2_mult160", model="upp16k", shape=(2, 4800), n_steps=3, seed=13, kwargs={}),
    dict(name="upp16k_3d", model="upp16k", shape=(2, 1, 3001), n_steps=2, seed=14, kwargs={}),
    dict(name="upp16k_1d_keeprms", model="upp16k", shape=(5000,), n_steps=3, seed=15,
         kwargs={"keep_rms": True}),
    dict(name="upp16k_eps", model="upp16k", shape=(1, 4000), n_steps=5, seed=16,
         kwargs={"epsilon": 2.0}),
    # aux-signal / warm-start path (universe.py:317-331, universe_gan.py:117-126,145-149: AliasFreeSnake)
    dict(name="upp16k_aux", model="upp16k", shape=(2, 3000), n_steps=2, seed=17,
         kwargs={"use_aux_signal": True}),
    dict(name="upp16k_warm", model="upp16k", shape=(1, 4000), n_steps=5, seed=18,
         kwargs={"warm_start": 2}),
    # ensembles (universe.py:261-264, 359-368; utils/stats.py:22-66)
    dict(name="upp16k_ens_median", model="upp16k", shape=(2, 2000), n_steps=2, seed=19,
         kwargs={"ensemble": 3, "ensemble_stat": "median"}),
    dict(name="upp16k_ens_sigmed", model="upp16k", shape=(2, 2000), n_steps=2, seed=19,
         kwargs={"ensemble": 3, "ensemble_stat": "signal_median"}),
    dict(name="upp16k_ens_mean", model="upp16k", shape=(1, 2000), n_steps=2, seed=20,
         kwargs={"ensemble": 2, "ensemble_stat": "mean"}),
    # sampler debugging with an oracle score (universe.py:276-300): the network is never called;
    # at +200 dB "score SNR" the torch.randn perturbation of the score is below fp32 resolution
    dict(name="upp16k_target", model="upp16k", shape=(2, 1, 2500), n_steps=6, seed=22,
         target_seed=23, kwargs={"fake_score_snr": 200.0}),
    dict(name="orig16k_short", model="orig16k", shape=(1, 6000), n_steps=3, seed=21, kwargs={}),
    dict(name="upp24k_short", model="upp24k", shape=(1, 7000), n_steps=3, seed=31, kwargs={}),
]

# module-level forwards on lengths that are NOT multiples of tot_ds: exercises the
# pad / crop bookkeeping of blocks.py:205-227,360-372 and score.py:104-127,294
NET_CASES = [
    dict(name="upp16k_net_odd", model="upp16k", B=2, T=5003, seed=41, sigmas=[0.3, 2.0]),
    dict(name="upp16k_net_even", model="upp16k", B=1, T=3200, seed=42, sigmas=[0.01]),
    dict(name="orig16k_net_odd", model="orig16k", B=1, T=3001, seed=43, sigmas=[1.1]),
    dict(name="upp24k_net_odd", model="upp24k", B=1, T=4001, seed=44, sigmas=[0.5]),
]


# UniverseLoRA.forward (networks/universe/lora.py:298-392): LoRA adapters (rank 8) on every Conv1d /
# ConvTranspose1d / Linear of both networks with non-zero factors; full sampler and partial diffusion
# (lora.py:231-296: per-clip random final time, no padding, no post-processing)
LORA_CASES = [
    dict(name="upp16k_lora_full", model="upp16k", shape=(2, 4000), n_steps=3, seed=51, partial=False),
    dict(name="upp16k_lora_partial", model="upp16k", shape=(2, 4800), n_steps=4, seed=52, partial=True),
]
LORA_RANK = 8
LORA_SEED = 9


def noise_rows(case):
    """Rows of the diffusion-noise tensors of an enhance case: batch x ensemble."""
    shape = tuple(case["shape"])
    b = 1 if len(shape) == 1 else shape[0]
    return b * (case["kwargs"].get("ensemble") or 1)


def case_kwargs(case):
    """enhance() keyword arguments of a case, with the synthetic ``target`` tensor materialised."""
    from detweights import det_audio
    kw = dict(case["kwargs"])
    if case.get("target_seed") is not None:
        kw["target"] = det_audio(tuple(case["shape"]), case["target_seed"])
    return kw
