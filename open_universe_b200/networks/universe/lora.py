"""``UniverseLoRA`` -- the LoRA fine-tuning wrapper of the reference (networks/universe/lora.py), forward
side only.

What is built: the constructor's model surgery (EMA weights copied in, loss modules dropped, weight norm
removed, LoRA adapters injected, requires_grad bookkeeping -- lora.py:62-165), ``enhance`` (positional
delegation, :200-229), ``forward`` (:298-392) and ``partial_diffusion`` (:231-296: a few sampler steps
with a per-clip random final time) on the CUDA path.  The adapted weights are merged at load time
(``engine.fold.effective_weight``), so the adapted network runs through the same kernels as the plain
one and the packed weights follow in-place updates of the LoRA factors (``runtime.weights_version``).
What is not: gradients.  The kernels have no backward, so outputs carry no ``grad_fn``
(``n_steps_backprop`` is accepted and ignored); ``training_step`` raises.
"""
import logging
from pathlib import Path
from typing import Optional

import torch

from ... import lora
from ...config import to_config
from .universe import randn  # noqa: F401  (module-level noise hook, as upstream)
from . import universe as _universe

log = logging.getLogger(__name__)


class UniverseLoRA(torch.nn.Module):
    def __init__(self, model, fs, losses=None, training=None, validation=None, optimizer=None,
                 scheduler=None, grad_clipper=None, diffusion=None, n_steps_backprop=1, use_lora=True,
                 use_lora_score=True, use_lora_condition=True, lora_rank=16, lora_alpha=None,
                 lora_train_biases=True, lora_train_names=[], use_hifigan_loss=False,
                 use_partial_diffusion=False, partial_diffusion_random_steps=False,
                 weight_hifigan_loss=1.0):
        super().__init__()
        if isinstance(model, (str, Path)):
            from ...inference_utils import load_model
            log.info(f"Loading pre-trained model from {model}")
            model = load_model(str(model))
        model.train()  # the loader puts the model in eval mode (lora.py:67)
        self.model = model
        self.fs = model.fs
        self.normalization_norm = model.normalization_norm
        self.normalization_kwargs = model.normalization_kwargs
        self.losses_kwargs, self.train_kwargs = to_config(losses), to_config(training)
        self.val_kwargs = to_config(validation)
        if n_steps_backprop < 1:
            raise ValueError("n_steps_backprop should be at least 1")
        if fs != self.model.fs:
            raise ValueError("The model fs should be the same as the input fs")
        self.n_steps_backprop = n_steps_backprop
        if diffusion is None:
            self.diff_kwargs = to_config({"n_steps": 8, "epsilon": 1.3})
        else:
            self.diff_kwargs = to_config(dict(diffusion) if isinstance(diffusion, dict) else diffusion)
        if use_hifigan_loss:
            raise NotImplementedError("the HiFi-GAN discriminators are training-only (not built)")
        self.use_hifigan_loss = False
        self.use_partial_diffusion = use_partial_diffusion
        self.partial_diffusion_random_steps = partial_diffusion_random_steps
        self.weight_hifigan_loss = weight_hifigan_loss
        self._fix_model()
        self.use_lora, self.lora_rank, self.lora_alpha = use_lora, lora_rank, lora_alpha
        if use_lora:
            if use_lora_score:
                lora.inject(self.model.get_score_model(), self.lora_rank, self.lora_alpha)
            if use_lora_condition:
                lora.inject(self.model.condition_model, self.lora_rank, self.lora_alpha)
            lora.freeze_parameters_except_lora_and_bias(self.model, train_biases=lora_train_biases,
                                                        train_names=lora_train_names)
        self.ema = None            # EMA of the trainable parameters is a training feature
        self._error_loading_ema = False
        self.rng = None

    def trainable_parameters(self):
        for p in self.parameters():
            if p.requires_grad:
                yield p

    def _fix_model(self):
        """lora.py:141-160: EMA weights become the weights, only the two networks are kept, weight norm is
        folded into plain weights."""
        if self.model.ema is not None:
            self.model.ema.copy_to(self.model.model_parameters())
            self.model.ema = None
        keep = ["score_model", "condition_model", "_edm_model"]
        for name, _ in list(self.model.named_children()):
            if name not in keep:
                # an attribute set to None keeps the inference code paths that test for it working
                setattr(self.model, name, None)
        self.model.remove_weight_norm()

    # ------------------------------------------------------------------ inference
    def enhance(self, mix, n_steps: Optional[int] = None, epsilon: Optional[float] = None,
                target: Optional[torch.Tensor] = None, fake_score_snr: Optional[float] = None,
                rng: Optional[torch.Generator] = None, use_aux_signal: Optional[bool] = False,
                keep_rms: Optional[bool] = False, ensemble: Optional[int] = None,
                ensemble_stat: Optional[str] = "median", warm_start: Optional[int] = None) -> torch.Tensor:
        if rng is None and getattr(self, "rng", None) is not None:
            rng = self.rng
        return self.model.enhance(mix, n_steps, epsilon, target, fake_score_snr, rng, use_aux_signal,
                                  keep_rms, ensemble, ensemble_stat, warm_start)

    def partial_diffusion(self, mix, rng=None, t_final=None):
        """lora.py:231-296 -- ``n_steps`` sampler steps from t = 1 to a per-clip final time drawn uniformly
        in [0, 1) (``t_final`` injects it for tests), WITHOUT padding or post-processing.  mix: (B, 1, T)."""
        n_steps = self.diff_kwargs.n_steps
        epsilon = self.diff_kwargs.epsilon
        if self.partial_diffusion_random_steps:
            n_steps = torch.randint(low=2, high=n_steps + 1, size=(1,), generator=rng,
                                    device=rng.device if rng is not None else "cpu").item()
        if t_final is None:
            t_final = mix.new_zeros(mix.shape[0]).uniform_(0, 1)
        delta_t = (1.0 - t_final) / (n_steps - 1)
        (mix, _), *_ = self.model.normalize_batch((mix, None))
        mix_wav = mix
        d = self.model.diff_kwargs
        gamma = (d.sigma_max / d.sigma_min) ** -delta_t
        eta = 1 - gamma**epsilon
        beta = torch.sqrt(1 - gamma ** (2 * (epsilon - 1.0)))
        time = mix.new_ones(mix.shape[0])
        sigma = self.model.get_std_dev(time)
        with torch.no_grad():
            cond, _, _ = self.model.condition_model(mix, x_wav=mix_wav, train=True)
            x = _universe.randn(mix, sigma, rng=rng)
            for n in range(0, n_steps - 1):
                score = self.model.score_model(x, sigma, cond)
                time = time - delta_t
                sigma_next = self.model.get_std_dev(time)
                z = _universe.randn(x, sigma_next, rng=rng)
                x = x + sigma[..., None, None] ** 2 * eta[..., None, None] * score + beta[..., None, None] * z
                sigma = sigma_next
            score = self.model.score_model(x, sigma, cond)
            x = x + sigma[:, None, None] ** 2 * score
        return x

    def forward(self, mix, n_steps: Optional[int] = None, epsilon: Optional[float] = None,
                rng: Optional[torch.Generator] = None, keep_rms: Optional[bool] = False) -> torch.Tensor:
        """lora.py:298-392.  The full-diffusion branch is the sampler of ``Universe.enhance`` (same pad /
        normalise / loop / unpad / limiter, no ensemble): it runs through the captured sampler loop."""
        x_ndim = mix.ndim
        if x_ndim > 3:
            raise ValueError("The input should have at most 3 dimensions")
        if not self.use_partial_diffusion:
            return self.model.enhance(mix, n_steps=n_steps, epsilon=epsilon, rng=rng, keep_rms=keep_rms)
        if x_ndim == 1:
            mix = mix[None, None, :]
        elif x_ndim == 2:
            mix = mix[:, None, :]
        x = self.partial_diffusion(mix, rng=rng)
        if x_ndim == 1:
            x = x[0, 0]
        elif x_ndim == 2:
            x = x[:, 0, :]
        return x

    def training_step(self, *args, **kwargs):
        raise NotImplementedError("fine-tuning needs backward kernels (SURVEY.md section 8(f) item 4)")
