"""``Universe`` -- the UNIVERSE / UNIVERSE++ model class with the B200-native ``enhance()``.

Drop-in for the inference surface of the reference's ``networks/universe/universe.py``:
constructor keywords (= Hydra YAML keys), attribute names, ``state_dict`` / EMA layout,
``pad`` / ``unpad`` / ``normalize_batch`` / ``get_std_dev`` / ``score_model`` /
``condition_model`` and the type-annotated ``enhance()`` signature (universe.py:231-244) that
``inference_utils.add_enhance_arguments`` reflects into argparse.

``enhance()`` (universe.py:231-375) here is a thin host driver: per call it launches
  1 x ``ou_pad_normalize``                         (pad + normalize_batch, :219-223, :272)
  1 x conditioner program                          (:314-316)
  2-5 launches for the sigma embedding / FiLM table of ALL steps, 5 hoisted 1x1 projections
  N x score program (~45 launches each; the last one fuses EDM mix + SDE update, :334-343)
  1 x ``ou_unpad_limit``                           (:349-357)
Training / validation / Lightning hooks are out of scope.
"""
import itertools
import logging
import math
from typing import Optional

import torch

from ... import utils
from ...config import Config, instantiate, to_config
from ...engine import lib, runtime
from ...layers.dyn_range_comp import IdentityTransform
from .blocks import remove_weight_norm

log = logging.getLogger(__name__)


def randn(x, sigma, rng=None):
    """Diffusion noise in the reference's draw order and scaling (universe.py:39-41).  Kept as a
    module-level function so that tests can inject pre-drawn noise exactly as with the reference."""
    noise = torch.randn(x.shape, dtype=x.dtype, device=x.device, generator=rng)
    return noise * sigma[:, None, None]


_default_randn = randn


def _draw_step_noise(out, sigma_next, rng):
    """Noise of one sampler step into the persistent buffer ``out`` (B, 1, T).  With the stock
    ``randn`` the values are the very ones ``randn(x, sigma_next)`` would draw (same generator
    consumption), but written in place and left at unit variance -- the caller folds sigma into the
    update coefficient -- which saves two 16 MB passes per step; a patched ``randn`` (test noise
    injection, batch-sharded draws) is honoured as is.  Returns True if ``out`` is unit-variance."""
    if randn is _default_randn:
        torch.randn(out.shape, dtype=out.dtype, device=out.device, generator=rng, out=out)
        return True
    out.copy_(randn(out, sigma_next, rng=rng))
    return False


class Universe(torch.nn.Module):
    def __init__(self, fs, normalization_norm, score_model, condition_model, diffusion, losses,
                 training, validation, optimizer, scheduler, grad_clipper, transform=None,
                 normalization_kwargs={}, with_noise_target=False, detach_cond=False, edm=None):
        super().__init__()
        score_model, condition_model = to_config(score_model), to_config(condition_model)
        self.fs = fs
        self.normalization_norm = normalization_norm
        self.normalization_kwargs = to_config(dict(normalization_kwargs))
        self.with_noise_target = with_noise_target
        self.detach_cond = detach_cond
        self.opt_kwargs = optimizer
        self.schedule_kwargs = scheduler
        self.grad_clip_kwargs = grad_clipper
        self.diff_kwargs = to_config(diffusion)
        self.losses_kwargs = to_config(losses)
        self.val_kwargs = to_config(validation)
        self.train_kwargs = to_config(training)

        if edm is not None:
            # EDM parameterisation (universe.py:80-92): the network is wrapped, so ``score_model``
            # is a bound method and the parameters live under ``_edm_model``
            self.edm_kwargs = to_config(edm)
            self._edm_model = instantiate(score_model, _recursive_=False)
            self.score_model = self._edm_score_wrapper
            self.with_edm = True
        else:
            self.score_model = instantiate(score_model, _recursive_=False)
            self.with_edm = False
        self.condition_model = instantiate(condition_model, _recursive_=False)
        self.n_channels = score_model.get("n_channels", 32)
        rate_factors = score_model.get("rate_factors", [2, 4, 4, 5])
        self.n_stages = len(rate_factors)
        self.latent_n_channels = 2**self.n_stages * self.n_channels
        self.tot_ds = math.prod(rate_factors)

        self.init_losses(score_model, condition_model, self.losses_kwargs, self.train_kwargs)
        self.denormalize_batch = utils.denormalize_batch
        self.transform = IdentityTransform()
        if transform is not None:
            raise NotImplementedError("only the identity transform is used by shipped configs "
                                      "(SURVEY.md section 2 #9)")

        self.ema_decay = getattr(self.train_kwargs, "ema_decay", 0.0) if self.train_kwargs else 0.0
        self._error_loading_ema = False
        if self.ema_decay and self.ema_decay > 0.0:
            self.ema = utils.ExponentialMovingAverage(self.model_parameters(), decay=self.ema_decay)
        else:
            self.ema = None

    # ------------------------------------------------------------------ structure / bookkeeping
    def init_losses(self, score_model, condition_model, losses, training):
        """Training losses (MDN heads, universe.py:135-167) are not instantiated: their
        ``loss_*`` checkpoint keys are skipped by ``inference_utils.load_model``."""

    def model_parameters(self):
        return itertools.chain(self.get_score_model().parameters(),
                               self.condition_model.parameters())

    def remove_weight_norm(self):
        remove_weight_norm(self)

    def get_score_model(self):
        return self._edm_model if self.with_edm else self.score_model

    def normalize_batch(self, batch, norm=None):
        if norm is None:
            norm = self.normalization_norm
        return utils.normalize_batch(batch, norm=norm, **self.normalization_kwargs)

    def pad(self, x, pad=None):
        """Centred zero padding to a multiple of the total down-sampling factor; a FULL extra
        ``tot_ds`` when the length already is a multiple (universe.py:219-223)."""
        if pad is None:
            pad = self.tot_ds - x.shape[-1] % self.tot_ds
        x = torch.nn.functional.pad(x, (pad // 2, pad - pad // 2))
        return x, pad

    def unpad(self, x, pad):
        return x[..., pad // 2: -(pad - pad // 2)]

    def aux_to_wav(self, y_aux):
        return y_aux

    def _aux_to_wav_blocked(self, y_blocked):
        """The conditioner's signal estimate (blocked bf16) -> (B, C, T) fp32 waveform(s)."""
        return self.aux_to_wav(runtime.unpack_blocked(y_blocked))

    def get_std_dev(self, time):
        if self.diff_kwargs.schedule == "geometric":
            s_min = self.diff_kwargs.sigma_min
            s_max = self.diff_kwargs.sigma_max
            return s_min * (s_max / s_min) ** time
        raise NotImplementedError()

    # ------------------------------------------------------------------ EDM wrapper
    def _sigma_data(self):
        level_db = self.edm_kwargs.get("data_level_db",
                                       self.normalization_kwargs.get("level_db", 0.0))
        return 10.0 ** (level_db / 20.0)

    def _get_edm_weights(self, sigma):
        """universe.py:175-189."""
        sigma_data = self._sigma_data()
        sigma_norm = (sigma**2 + sigma_data**2) ** 0.5
        return {"skip": sigma_data**2 / (sigma**2 + sigma_data**2), "in": 1.0 / sigma_norm,
                "out": sigma * sigma_data / sigma_norm, "noise": self.edm_kwargs.noise}

    @runtime.on_tensor_device
    def _edm_score_wrapper(self, x, sigma, cond, with_speech_est=False):
        """score = (w_skip x + w_out net(w_in x, c_noise sigma) - x) / sigma^2 (universe.py:197-209),
        with the affine combination evaluated inside the network's last kernel."""
        runtime.require_cuda(x, sigma, *cond)
        net = self._edm_model
        b, _, t = x.shape
        w = self._get_edm_weights(sigma.double())
        r = runtime.get_score_runner(net, b, t, x.device)
        with torch.no_grad():
            r.set_sigmas((w["noise"] * sigma.double()).float().reshape(-1))
            r.set_cond([runtime.pack_blocked(c) for c in cond])
            s2 = sigma.double() ** 2
            coef = torch.stack([(w["skip"] - 1.0) / s2, w["out"] / s2, torch.zeros_like(s2)],
                               dim=1).float().contiguous()
            score = torch.empty(b, 1, t, dtype=torch.float32, device=x.device)
            xin = x.contiguous().float()
            r.step(xin, 0, True, in_scale=w["in"].float().contiguous(), coef=coef, xout=score)
            if with_speech_est:
                return score, xin + utils.pad_dim_right(sigma.float(), xin) ** 2 * score
        return score

    def forward(self, xt, sigma, cond):
        return self.score_model(xt, sigma, cond)

    # ------------------------------------------------------------------ the hot path
    def _sampler_tables(self, n_steps, epsilon, like):
        """Per-step scalars of the reverse SDE (universe.py:301-311) and of the EDM wrapper
        (universe.py:175-209), reduced to the affine form the fused last kernel applies:
            x <- ca * x + cb * net + cc * z         (z already scaled by sigma_next, as upstream)
        Returns sigma (N,), net_sigma (N,), in_scale (N,) and coef (N, 3), all fp32 on device."""
        d = self.diff_kwargs
        delta_t = 1.0 / (n_steps - 1)
        gamma = (d.sigma_max / d.sigma_min) ** -delta_t
        eta = 1 - gamma**epsilon
        beta = math.sqrt(1 - gamma ** (2 * (epsilon - 1.0)))
        time = torch.linspace(0, 1, n_steps).type_as(like).flip(dims=[0])
        sigma = self.get_std_dev(time)                      # fp32, same arithmetic as upstream
        s = sigma.double()
        last = torch.zeros_like(s)
        last[-1] = 1.0
        eta_n = eta * (1 - last) + last                     # the final step uses eta = 1, no noise
        if self.with_edm:
            sd = self._sigma_data()
            norm = (s**2 + sd**2).sqrt()
            w_skip, w_out = sd**2 / (s**2 + sd**2), s * sd / norm
            ca = 1 + eta_n * (w_skip - 1)
            cb = eta_n * w_out
            in_scale = 1.0 / norm
            net_sigma = self.edm_kwargs.noise * s
        else:
            ca = torch.ones_like(s)
            cb = s**2 * eta_n
            in_scale = torch.ones_like(s)
            net_sigma = s
        cc = beta * (1 - last)
        coef = torch.stack([ca, cb, cc], dim=1)
        return sigma, net_sigma.float(), in_scale.float(), coef.float(), (eta, beta)

    @runtime.on_tensor_device
    def enhance(
        self,
        mix,
        n_steps: Optional[int] = None,
        epsilon: Optional[float] = None,
        target: Optional[torch.Tensor] = None,
        fake_score_snr: Optional[float] = None,
        rng: Optional[torch.Generator] = None,
        use_aux_signal: Optional[bool] = False,
        keep_rms: Optional[bool] = False,
        ensemble: Optional[int] = None,
        ensemble_stat: Optional[str] = "median",
        warm_start: Optional[int] = None,
    ) -> torch.Tensor:
        if epsilon is None:
            epsilon = self.diff_kwargs.epsilon
        if n_steps is None:
            n_steps = self.diff_kwargs.n_steps
        if n_steps < 2:
            raise ValueError("n_steps must be at least 2")
        if warm_start is not None and not 0 <= warm_start < n_steps:
            raise ValueError("warm_start must be a step index in [0, n_steps)")
        if ensemble is not None and ensemble_stat not in ("mean", "median", "signal_median"):
            raise NotImplementedError()

        x_ndim = mix.ndim
        if x_ndim == 1:
            mix = mix[None, None, :]
        elif x_ndim == 2:
            mix = mix[:, None, :]
        elif x_ndim > 3:
            raise ValueError("The input should have at most 3 dimensions")
        runtime.require_cuda(mix)
        p_dev = next(self.parameters()).device
        if p_dev != mix.device:
            raise ValueError(f"model weights are on {p_dev} but the input is on {mix.device}")
        if mix.shape[1] != 1:
            raise NotImplementedError("multi-channel clips: pass channels as batch rows (B, T)")
        L = lib.load()
        dev = mix.device
        mix = mix.contiguous().float()
        mix_rms = mix.square().mean(dim=(-2, -1)).sqrt().contiguous() if keep_rms else None

        mix_shape = mix.shape
        if ensemble is not None:
            mix = torch.stack([mix] * ensemble, dim=0).view((-1,) + mix_shape[1:]).contiguous()
            if mix_rms is not None:
                mix_rms = mix_rms.repeat(ensemble).contiguous()
        if target is not None:
            if ensemble is not None:
                raise NotImplementedError("target (oracle score) and ensemble do not combine upstream")
            runtime.require_cuda(target)
            target = target.reshape(mix.shape).contiguous().float()
        B, _, mix_len = mix.shape
        pad = self.tot_ds - mix_len % self.tot_ds
        t_pad = mix_len + pad
        level = 10 ** (self.normalization_kwargs.get("level_db", 0.0) / 20.0)
        if self.normalization_norm not in (2, "2"):
            raise NotImplementedError("only normalization_norm=2 is used by shipped configs")

        with torch.no_grad():
            # pad + normalize_batch fused (universe.py:267-272)
            mixn = torch.empty(B, 1, t_pad, dtype=torch.float32, device=dev)
            stats = torch.empty(B, 2, dtype=torch.float32, device=dev)
            lib.check(L.ou_pad_normalize(runtime._ptr(mix), runtime._ptr(mixn), runtime._ptr(stats), B,
                                         mix_len, t_pad, pad // 2, level, runtime._stream()))
            if target is not None:
                # utils/norm.py:72-86: own statistics with ref='both', the mixture's otherwise
                tgt = torch.empty_like(mixn)
                if self.normalization_kwargs.get("ref", "noisy") == "both":
                    lib.check(L.ou_pad_normalize(runtime._ptr(target), runtime._ptr(tgt), None, B,
                                                 mix_len, t_pad, pad // 2, level, runtime._stream()))
                else:
                    tp = torch.nn.functional.pad(target, (pad // 2, pad - pad // 2))
                    tgt = (tp - stats[:, 0, None, None]) / stats[:, 1, None, None]
            sigma, net_sigma, in_scale, coef, _ = self._sampler_tables(n_steps, epsilon, mixn)
            sigma_b = torch.broadcast_to(sigma[None, :], (B, n_steps))
            in_scale_b = in_scale[:, None].expand(n_steps, B).contiguous()
            coef_b = coef[:, None, :].expand(n_steps, B, 3).contiguous()

            # conditioning: once per call (universe.py:314-316)
            need_aux = bool(use_aux_signal) or warm_start is not None
            cr = runtime.get_conditioner_runner(self.condition_model, B, t_pad, dev, need_aux)
            cond, aux_signal, _ = cr.run(mixn, mixn)
            sig = None
            if need_aux:
                # the conditioner's signal estimate -> waveform (universe.py:317-319, 328)
                sig = self._aux_to_wav_blocked(aux_signal)
                if sig.shape != (B, 1, t_pad):
                    raise ValueError("aux_to_wav must return one waveform per clip, got "
                                     f"{tuple(sig.shape)}")
            if use_aux_signal:
                x = sig
            elif target is not None:
                x = self._oracle_score_sampler(mixn, tgt, sig, sigma_b, n_steps, epsilon, warm_start,
                                               fake_score_snr, rng)
            else:
                sr = runtime.get_score_runner(self.get_score_model(), B, t_pad, dev, pipelined=True)
                sr.set_sigmas(net_sigma)
                sr.set_cond(cond)
                # the N-step loop (universe.py:334-343) replays a CUDA graph over persistent
                # buffers; noise is drawn with torch.randn in the reference's order, one call per
                # step.  warm_start = n0 starts from the signal estimate at noise level sigma[n0]
                # (universe.py:326-331)
                n_start = 0 if warm_start is None else int(warm_start)
                loop = runtime.get_sampler_loop(sr, n_steps, n_start)
                loop.in_scale.copy_(in_scale_b)
                if warm_start is None:
                    loop.x.copy_(randn(mixn, sigma_b[:, 0], rng=rng))
                else:
                    loop.x.copy_(sig + randn(sig, sigma_b[:, n_start], rng=rng))
                unit = [_draw_step_noise(loop.noise[n], sigma_b[:, n + 1], rng)
                        for n in range(n_start, n_steps - 1)]
                if unit and all(unit):
                    # z = sigma_{n+1} * noise: the scale goes into the coefficient cc of step n
                    coef_b = coef_b.clone()
                    coef_b[:-1, :, 2] *= sigma[1:, None]
                elif any(unit):
                    raise RuntimeError("randn was patched while enhance() was drawing noise")
                loop.coef.copy_(coef_b)
                loop.run()
                x = loop.x

            # unpad, keep_rms, peak limiter (universe.py:346-357)
            out = torch.empty(B, 1, mix_len, dtype=torch.float32, device=dev)
            lib.check(L.ou_unpad_limit(runtime._ptr(x), runtime._ptr(mix_rms), runtime._ptr(out), B,
                                       t_pad, pad // 2, mix_len, mix_len, runtime._stream()))
            x = out

        if ensemble is not None:
            x = x.view((-1,) + mix_shape)
            if ensemble_stat == "mean":
                x = x.mean(dim=0)
            elif ensemble_stat == "median":
                x = x.median(dim=0).values
            elif ensemble_stat == "signal_median":
                x = utils.signal_median(x)
            else:
                raise NotImplementedError()
        if x_ndim == 1:
            x = x[0, 0]
        elif x_ndim == 2:
            x = x[:, 0, :]
        return x

    def _oracle_score_sampler(self, mixn, target, sig, sigma, n_steps, epsilon, warm_start,
                              fake_score_snr, rng):
        """The sampler driven by a "perfect" score computed from a known target plus white noise at
        ``fake_score_snr`` dB (universe.py:276-300, 322-343).  A debugging aid for the sampler
        itself: the score network is never evaluated, so this is a dozen element-wise torch ops on
        (B, 1, T) signals per step, not a kernel path."""
        d = self.diff_kwargs
        delta_t = 1.0 / (n_steps - 1)
        gamma = (d.sigma_max / d.sigma_min) ** -delta_t
        eta = 1 - gamma**epsilon
        beta = math.sqrt(1 - gamma ** (2 * (epsilon - 1.0)))
        score_snr = 5.0 if fake_score_snr is None else fake_score_snr

        def score(x, s):
            true_score = -(x - target) / s[:, None, None] ** 2
            noise_rms = (true_score**2).mean().sqrt() * 10 ** (-score_snr / 20.0)
            nz = torch.randn(true_score.shape, dtype=true_score.dtype, device=true_score.device,
                             generator=rng)
            return true_score + nz * noise_rms

        if warm_start is None:
            x, n_start = randn(mixn, sigma[:, 0], rng=rng), 0
        else:
            x, n_start = sig + randn(sig, sigma[:, warm_start], rng=rng), int(warm_start)
        for n in range(n_start, n_steps - 1):
            sc = score(x, sigma[:, n])
            z = randn(x, sigma[:, n + 1], rng=rng)
            x = x + sigma[:, n, None, None] ** 2 * eta * sc + beta * z
        return x + sigma[:, -1, None, None] ** 2 * score(x, sigma[:, -1])

    # ------------------------------------------------------------------ validation (enhancement half)
    def on_validation_epoch_start(self):
        """universe.py:588-604: counters and the de-randomised generator of the validation loop."""
        self.n_batches_est_done = 0
        val = self.val_kwargs or {}
        self.num_tb_samples = val.get("num_tb_samples", 0)
        dev = next(self.parameters()).device
        self.rng = torch.Generator(device=dev)
        self.rng.manual_seed(682479040)

    def on_validation_epoch_end(self):
        self.rng = None

    def validation_step(self, batch, batch_idx, dataset_i=0):
        """The enhancement half of the reference's validation step (universe.py:651-660, 704-719):
        ``est = enhance(mix, rng=self.rng)`` on the un-normalised batch, limited to
        ``validation.max_enh_batches`` batches per epoch, then every configured ``enh_losses`` metric on
        (est, target).  The log-spectral distance runs as a CUDA kernel (``metrics.LogSpectralDistance``,
        always reported as ``"lsd"``); the score-matching loss bins of the first half (universe.py:606-650)
        need the training forward (MDN losses, autograd) and are not built.  Returns {name: value} -- there
        is no Lightning logger here -- or None once the batch budget is used up."""
        if getattr(self, "rng", None) is None:
            self.on_validation_epoch_start()
        mix, target = batch[:2]
        val = self.val_kwargs or {}
        max_batches = val.get("max_enh_batches", None)
        if max_batches is not None and self.n_batches_est_done >= max_batches:
            return None
        self.n_batches_est_done += 1
        with torch.no_grad():
            est = self.enhance(mix, rng=self.rng)
            from ...metrics import LogSpectralDistance
            if getattr(self, "_lsd_metric", None) is None:
                self._lsd_metric = LogSpectralDistance().to(est.device)
            out = {"lsd": self._lsd_metric(est, target.reshape(est.shape))}
            for name, loss in (getattr(self, "enh_losses", None) or {}).items():
                metric = loss(est, target)
                if not isinstance(metric, dict):
                    metric = {"": metric}
                for sub, value in metric.items():
                    out[name + sub] = value
        out["est"] = est
        return out

    # ------------------------------------------------------------------ EMA weight swap
    def train(self, mode=True, no_ema=False):
        """eval() copies the EMA shadow weights into the live parameters, train() restores them
        (universe.py:841-865).  Packed device weights are re-derived lazily (runtime cache keyed on
        the parameters' version counters)."""
        res = super().train(mode)
        if getattr(self, "ema", None) is None:
            return res
        if not self._error_loading_ema:
            if mode is False and not no_ema:
                self.ema.store(self.model_parameters())
                self.ema.copy_to(self.model_parameters())
            elif self.ema.collected_params is not None:
                self.ema.restore(self.model_parameters())
        return res

    def eval(self, no_ema=False):
        return self.train(False, no_ema=no_ema)

    def to(self, *args, **kwargs):
        res = super().to(*args, **kwargs)
        if self.ema is not None:
            p = next(self.parameters())
            self.ema.to(device=p.device)
        return res

    def on_load_checkpoint(self, checkpoint):
        ema = checkpoint.get("ema", None)
        if self.ema is not None:
            if ema is not None:
                self.ema.load_state_dict(ema)
            else:
                self._error_loading_ema = True
                log.warning("EMA state_dict not found in checkpoint!")

    def on_save_checkpoint(self, checkpoint):
        if self.ema is not None:
            checkpoint["ema"] = self.ema.state_dict()
