"""The UNIVERSE++ and UNIVERSE models (Hydra ``_target_`` namespace of the reference:
``open_universe.networks.universe.{Universe, UniverseGAN, ScoreNetwork, ConditionerNetwork}``)."""
from .condition import ConditionerNetwork
from .lora import UniverseLoRA
from .score import ScoreNetwork
from .universe import Universe
from .universe_gan import UniverseGAN

__all__ = ["ConditionerNetwork", "ScoreNetwork", "Universe", "UniverseGAN", "UniverseLoRA"]
