"""``UniverseGAN`` -- the UNIVERSE++ model class (reference ``networks/universe/universe_gan.py``).

At inference UNIVERSE++ is the same sampler as UNIVERSE over a network with weight norm,
anti-aliased rate changes and the EDM wrapper; the "HiFi-GAN" part is a pair of training-only
discriminators (universe_gan.py:107-134) which are not instantiated here (their ``loss_mpd.*`` /
``loss_mrd.*`` checkpoint keys are skipped by the loader).  What remains inference-relevant:
``transform`` is forced to identity (universe_gan.py:91) and the ``signal_decoupling_layer``
(:117-126) used by ``aux_to_wav`` for ``warm_start`` / ``use_aux_signal``.
"""
import itertools

from ...config import to_config
from .blocks import PReLU_Conv
from .universe import Universe


class UniverseGAN(Universe):
    def __init__(self, fs, normalization_norm, score_model, condition_model, diffusion, losses,
                 training, validation, optimizer, scheduler, grad_clipper, transform=None,
                 normalization_kwargs={}, detach_cond=False, edm=None):
        super().__init__(fs, normalization_norm, score_model, condition_model, diffusion, losses,
                         training, validation, optimizer, scheduler, grad_clipper, transform=None,
                         normalization_kwargs=normalization_kwargs, detach_cond=detach_cond, edm=edm)
        self.automatic_optimization = False

    def init_losses(self, score_model, condition_model, losses, training):
        losses = to_config(losses or {})
        if losses.get("use_signal_decoupling", False):
            self.signal_decoupling_layer = PReLU_Conv(
                self.n_channels, 1, kernel_size=3, padding="same",
                act_type=losses.get("signal_decoupling_act", None))
        else:
            self.signal_decoupling_layer = None
        self.disc_freeze_step = losses.get("disc_freeze_step", 0)

    def model_parameters(self):
        params = itertools.chain(self.get_score_model().parameters(),
                                 self.condition_model.parameters())
        if self.signal_decoupling_layer is not None:
            params = itertools.chain(params, self.signal_decoupling_layer.parameters())
        return params

    def aux_to_wav(self, y_aux):
        if self.signal_decoupling_layer is not None:
            return self.signal_decoupling_layer(y_aux)
        return y_aux

    def _aux_to_wav_blocked(self, y_blocked):
        """enhance()-internal: the conditioner's signal output still in the blocked bf16 layout."""
        from ...engine import runtime
        if self.signal_decoupling_layer is not None:
            return runtime.prelu_conv_forward(self.signal_decoupling_layer, y_blocked, blocked=True)
        return super()._aux_to_wav_blocked(y_blocked)
