"""GPU kernels around the path: polyphase resampler (CLI) and log-spectral distance, against the
third-party routines the reference calls (torchaudio, run on the CPU here as the checker)."""
import pytest
import torch

from common import det_audio, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("fs,target", [(16000, 24000), (24000, 16000), (44100, 16000), (16000, 44100),
                                       (8000, 16000), (48000, 16000)])
@pytest.mark.parametrize("shape", [(2, 8000), (1, 12345)])
def test_resample_vs_torchaudio(fs, target, shape):
    torchaudio = pytest.importorskip("torchaudio")
    from open_universe_b200.utils.resample import resample
    x = det_audio(shape, 7, level=0.3)
    want = torchaudio.functional.resample(x, fs, target)
    got = resample(x.to(DEV), fs, target).cpu()
    assert got.shape == want.shape
    # fp32 accumulation over up to 475 taps; the filter bank is built in float64 here, in float32 by torchaudio
    assert rel_rms(got, want) < 2e-5


def test_resample_identity_and_leading_dims():
    from open_universe_b200.utils.resample import resample
    x = det_audio((2, 3, 1000), 8).to(DEV)
    assert resample(x, 16000, 16000) is x
    y = resample(x, 16000, 8000)
    assert y.shape == (2, 3, 500)


def torch_lsd(input, target, p=2, db=True, n_fft=400, hop=160, eps=1e-7, scale_invariant=False):
    """metrics/lsd.py:84-147 restated with torch.stft (what torchaudio.functional.spectrogram calls)."""
    window = torch.hann_window(n_fft, periodic=True, dtype=input.dtype)
    sf = (torch.sum(input * target, -1, keepdim=True) / (torch.sum(input**2, -1, keepdim=True) + eps)
          if scale_invariant else 1.0)

    def logspec(x):
        s = torch.stft(x, n_fft, hop_length=hop, win_length=n_fft, window=window, center=True,
                       pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
        s = s / window.pow(2.0).sum().sqrt()
        pw = s.abs().pow(2.0)
        return 10 * torch.log10(pw + eps) if db else torch.log(pw + eps)

    a, b = logspec(input), logspec(sf * target)
    denom = (b.shape[-1] * b.shape[-2]) ** (1 / p)
    return torch.norm(a - b, p=p, dim=(-2, -1)) / denom


@pytest.mark.parametrize("kw", [dict(), dict(p=1, db=False), dict(scale_invariant=True), dict(n_fft=512, hop=128),
                                dict(p=3)], ids=str)
def test_lsd_vs_torch(kw):
    from open_universe_b200.metrics import log_spectral_distance
    x = det_audio((3, 16000), 11, level=0.1)
    y = 0.7 * x + det_audio((3, 16000), 12, level=0.05)
    want = torch_lsd(x, y, **kw)
    kw2 = dict(kw)
    if "hop" in kw2:
        kw2["hop_length"] = kw2.pop("hop")
    got = log_spectral_distance(x.to(DEV), y.to(DEV), **kw2).cpu()
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=2e-4, atol=1e-4), (got, want)


def test_lsd_matches_torchaudio_spectrogram_and_module():
    torchaudio = pytest.importorskip("torchaudio")
    from open_universe_b200.metrics import LogSpectralDistance
    x = det_audio((2, 2, 8000), 13, level=0.1)
    y = det_audio((2, 2, 8000), 14, level=0.1)
    window = torch.hann_window(400, periodic=True)

    def logspec(v):
        s = torchaudio.functional.spectrogram(v, pad=0, win_length=400, window=window, n_fft=400, hop_length=160,
                                              power=2, normalized="window")
        return 10 * torch.log10(s + 1e-5)

    a, b = logspec(x), logspec(y)
    want = (torch.norm(a - b, p=2, dim=(-2, -1)) / (b.shape[-1] * b.shape[-2]) ** 0.5).mean()
    got = LogSpectralDistance().to(DEV)(x.to(DEV), y.to(DEV)).cpu()
    assert torch.allclose(got, want, rtol=2e-4)


def test_validation_step_enhancement_half():
    """Universe.validation_step: de-randomised generator (seed 682479040, universe.py:602-604), enhance()
    on the raw batch, LSD kernel on (est, target), batch budget ``validation.max_enh_batches``."""
    from open_universe_b200.config import builtin_config, instantiate
    torch.manual_seed(0)
    m = instantiate(builtin_config("universepp_16k").model, _recursive_=False)
    m.eval(no_ema=True)
    m = m.to(DEV)
    mix, target = det_audio((2, 4000), 21).to(DEV), det_audio((2, 4000), 22).to(DEV)
    m.on_validation_epoch_start()
    out = m.validation_step((mix, target), 0)
    want = m.enhance(mix, rng=torch.Generator(device=DEV).manual_seed(682479040))
    assert torch.equal(out["est"], want)
    ref = torch_lsd(want.cpu(), target.cpu(), eps=1e-5).mean()
    assert torch.allclose(out["lsd"].cpu(), ref, rtol=2e-4)
    m.val_kwargs["max_enh_batches"] = 1
    assert m.validation_step((mix, target), 1) is None
    m.on_validation_epoch_end()
    assert m.rng is None
