"""open_universe_b200 -- B200-native (sm_100a) implementation of open-universe's enhance() hot path.

Drop-in surface: ``networks.universe.{Universe, UniverseGAN, ScoreNetwork,
ConditionerNetwork}``, ``inference_utils.{load_model, add_enhance_arguments}``.  All device work
goes through the C-ABI library ``csrc/libou_b200.so`` (``include/ou_b200.h``); there is no CPU
fallback -- calling a forward without the CUDA library / a CUDA device raises.
"""
__version__ = "0.1.0"

from . import config  # noqa: F401
