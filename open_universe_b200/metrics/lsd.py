"""Log-spectral distance (reference metrics/lsd.py): same functional and module signatures, computed by
``ou_lsd`` (framing with reflect padding, windowed DFT, log-power difference and p-norm in one pass)."""
import torch

from ..engine import lib, runtime


@runtime.on_tensor_device
def log_spectral_distance(input, target, p=2, db=True, n_fft=400, hop_length=160, eps=1e-7, win_length=None,
                          window=None, pad=0, scale_invariant=False, **stft_kwargs):
    """lsd.py:26-147.  input / target: (..., T) CUDA tensors -> (...) fp32.  ``win_length`` must equal
    ``n_fft`` and ``pad`` be 0 (the reference's defaults and every call site)."""
    if win_length is None:
        win_length = n_fft
    if p is None or p <= 0:
        raise ValueError(f"p must be a positive number, but got p={p}")
    if win_length != n_fft or pad != 0 or stft_kwargs:
        raise NotImplementedError("only win_length == n_fft, pad == 0 and default STFT options are built")
    runtime.require_cuda(input, target)
    runtime.same_device(input, target)
    if input.shape != target.shape:
        raise ValueError("input and target must have the same shape")
    shape = input.shape
    T = shape[-1]
    x = input.reshape(-1, T).contiguous().float()
    y = target.reshape(-1, T).contiguous().float()
    dev = x.device
    if window is None:
        window = torch.hann_window(n_fft, periodic=True, dtype=torch.float32, device=dev)
    window = window.to(dev).contiguous().float()
    if window.numel() != n_fft:
        raise ValueError("window length must equal n_fft")
    B = x.shape[0]
    frames = 1 + T // hop_length
    partial = torch.empty(B, frames, dtype=torch.float32, device=dev)
    scale = torch.empty(B, dtype=torch.float32, device=dev)
    out = torch.empty(B, dtype=torch.float32, device=dev)
    wss = float(window.double().square().sum().item())
    lib.check(lib.load().ou_lsd(runtime._ptr(x), runtime._ptr(y), runtime._ptr(window), runtime._ptr(partial),
                                runtime._ptr(scale), runtime._ptr(out), B, T, n_fft, hop_length, frames,
                                float(p), 1 if db else 0, float(eps), wss, 1 if scale_invariant else 0,
                                runtime._stream()))
    return out.reshape(shape[:-1])


class LogSpectralDistance(torch.nn.Module):
    """lsd.py:150-215."""

    def __init__(self, p=2, db=True, n_fft=400, hop_length=160, win_length=None, window=None, pad=0,
                 eps=1e-5, reduction="mean", scale_invariant=False, **stft_kwargs):
        super().__init__()
        self.p, self.eps, self.n_fft, self.hop_length = p, eps, n_fft, hop_length
        self.win_length = win_length if win_length is not None else n_fft
        self.pad, self.stft_kwargs, self.db = pad, stft_kwargs, db
        self.reduction, self.scale_invariant = reduction, scale_invariant
        if reduction not in ("mean", "sum", "none"):
            raise ValueError("Only reduction=mean|sum|none are supported")
        if window is None:
            window = torch.hann_window(self.win_length, periodic=True)
        self.register_buffer("window", window)

    def forward(self, input, target):
        dist = log_spectral_distance(input, target, p=self.p, db=self.db, n_fft=self.n_fft,
                                     hop_length=self.hop_length, pad=self.pad, window=self.window,
                                     win_length=self.win_length, eps=self.eps,
                                     scale_invariant=self.scale_invariant, **self.stft_kwargs)
        if self.reduction == "mean":
            return dist.mean()
        if self.reduction == "sum":
            return dist.sum()
        return dist
