"""Host-side helpers mirroring the reference's ``open_universe/utils`` (inference-relevant part)."""
from .ema import ExponentialMovingAverage  # noqa: F401
from .norm import denormalize_batch, normalize_batch  # noqa: F401
from .stats import signal_median  # noqa: F401
from .torch_utils import count_parameters, pad_dim_right  # noqa: F401
