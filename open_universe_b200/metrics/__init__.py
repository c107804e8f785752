"""Post-path metrics with a CUDA kernel: the log-spectral distance.
The other metrics of the reference (PESQ, STOI, DNSMOS, ...) wrap third-party packages and are out of scope."""
from .lsd import LogSpectralDistance, log_spectral_distance

__all__ = ["LogSpectralDistance", "log_spectral_distance"]
