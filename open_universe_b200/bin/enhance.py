"""Inference CLI: enhance a file or a folder of audio files -- the caller of the hot path.

Drop-in for the reference's ``python -m open_universe.bin.enhance`` (``bin/enhance.py:83-192``): same
positional arguments, same ``--model / --hf-token / --model-strict / --seed / --device`` options, the
model's own ``enhance()`` options reflected into argparse by ``add_enhance_arguments``, the folder
structure of the input is kept.  Two additions that turn kernel throughput into end-user throughput
(SURVEY.md section 8(f) item 2):

* ``--batch-size N`` (default 1 = the reference's one-file-per-call behaviour, bit-compatible noise
  order): files whose channel rows have the same sample rate and length are enhanced together, up
  to N rows per ``enhance()`` call (every row of a batch is independent end to end, so batching
  changes nothing but the diffusion-noise draw order);
* host <-> device copies go through pinned buffers on a side stream so that reading / resampling
  the next group overlaps the current ``enhance()``.

Audio I/O uses ``torchaudio.load/save`` when a codec backend is installed and falls back to
``scipy.io.wavfile`` for WAV files otherwise.  There is no CPU inference path: without CUDA the
script stops with an error instead of silently running something else.
"""
import argparse
import sys
from collections import OrderedDict
from pathlib import Path

import torch

AUDIO_SUFFIXES = [".wav", ".mp3", ".flac"]


# ------------------------------------------------------------------------------------ audio I/O
def load_audio(path):
    """-> (channels, samples) float32 tensor in [-1, 1], sample rate."""
    try:
        import torchaudio
        audio, fs = torchaudio.load(str(path))
        return audio.float(), int(fs)
    except (ImportError, RuntimeError, OSError):
        if Path(path).suffix.lower() != ".wav":
            raise
    import numpy as np
    from scipy.io import wavfile
    fs, data = wavfile.read(str(path))
    if data.dtype.kind == "i":
        data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
    elif data.dtype.kind == "u":                      # 8-bit PCM is unsigned
        data = (data.astype(np.float32) - 128.0) / 128.0
    else:
        data = data.astype(np.float32)
    if data.ndim == 1:
        data = data[:, None]
    return torch.from_numpy(np.ascontiguousarray(data.T)), int(fs)


def save_audio(path, audio, fs):
    """audio: (channels, samples) float tensor on the CPU."""
    try:
        import torchaudio
        torchaudio.save(str(path), audio, fs)
        return
    except (ImportError, RuntimeError, OSError):
        if Path(path).suffix.lower() != ".wav":
            raise
    from scipy.io import wavfile
    wavfile.write(str(path), fs, audio.t().contiguous().numpy().astype("float32"))


def resample(audio, fs, target_fs):
    """Sample-rate conversion on the GPU (upstream: torchaudio.functional.resample, enhance.py:77-80): the
    polyphase windowed-sinc kernel ``ou_resample_poly`` with torchaudio's default filter design."""
    if fs != target_fs:
        from ..utils.resample import resample as gpu_resample
        audio = gpu_resample(audio, fs, target_fs)
    return audio


# ------------------------------------------------------------------------------------ file handling
def handle_help(argv):
    """Defer --help until the model's own arguments are known (reference enhance.py:36-58)."""
    if "--model" not in argv:
        return False
    for flag in ("--help", "-h"):
        if flag in argv:
            argv.remove(flag)
            return True
    return False


def find_files(path):
    if not path.is_dir():
        return [path], path.parent, False
    return sorted(p for p in path.rglob("*") if p.suffix in AUDIO_SUFFIXES), path, True


def output_path_for(path, rel_path, dir_proc, output):
    if dir_proc:
        out = output / path.relative_to(rel_path)
        out.parent.mkdir(exist_ok=True, parents=True)
        return out
    return output / path.name if output.is_dir() else output


MAX_OPEN_GROUPS = 64   # bound on files held in memory while waiting for partners of equal length


def group_files(infos, batch_size):
    """infos: iterable of (path, fs, channels, samples).  Yields lists of infos that share
    (fs, samples) with at most ``batch_size`` channel rows in total, in first-seen order; a file is
    never split across groups (a file with more rows than ``batch_size`` forms its own group)."""
    open_groups = OrderedDict()
    for info in infos:
        if len(open_groups) > MAX_OPEN_GROUPS:
            yield open_groups.popitem(last=False)[1][0]
        _, fs, ch, n = info
        key = (fs, n)
        cur = open_groups.get(key)
        if cur is not None and cur[1] + ch > batch_size:
            yield open_groups.pop(key)[0]
            cur = None
        if cur is None:
            cur = open_groups[key] = [[], 0]
        cur[0].append(info)
        cur[1] += ch
        if cur[1] >= batch_size:
            yield open_groups.pop(key)[0]
    for files, _ in open_groups.values():
        yield files


def build_parser():
    parser = argparse.ArgumentParser(description="Enhance a file or a directory of audio files")
    parser.add_argument("input", type=Path, help="Path to an audio file or a folder of audio files")
    parser.add_argument("output", type=Path,
                        help="Output path for the enhanced files. In the case of a folder, the "
                             "original structure is retained.")
    parser.add_argument("--model", type=str, default="line-corporation/open-universe:plusplus",
                        help="A checkpoint path or a model of the Hugging Face model zoo")
    parser.add_argument("--hf-token", type=str, help="Huggingface access token")
    parser.add_argument("--model-strict", action="store_true",
                        help="Use strict policy to load the model. Can help uncover problems.")
    parser.add_argument("--seed", type=int, default=1028282,
                        help="Set a deterministic seed to get reproducible results")
    parser.add_argument("--device", type=str, default="cuda:0", help="The CUDA device to use")
    parser.add_argument("--batch-size", type=int, default=1,
                        help="Enhance up to this many equally long channel rows per call "
                             "(1 = one file per call, as the reference)")
    return parser


def main(argv=None):
    from open_universe_b200 import inference_utils
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = build_parser()
    requires_help = handle_help(argv)
    args, _ = parser.parse_known_args(argv)
    if not args.device.startswith("cuda"):
        raise ValueError("open_universe_b200 runs on CUDA devices only (no CPU inference path); "
                         f"got --device {args.device}")
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available and there is no CPU fallback")
    device = torch.device(args.device)
    torch.cuda.set_device(device)
    model = inference_utils.load_model(args.model, device=device, strict=args.model_strict,
                                       hf_token=args.hf_token)
    rng = torch.Generator(device=device)
    rng.manual_seed(args.seed)
    inference_utils.add_enhance_arguments(model, parser)
    if requires_help:
        argv.append("--help")
    args = parser.parse_args(argv)
    enhance_kwargs = {}
    for group in parser._action_groups:
        if group.title == "enhance":
            enhance_kwargs = {a.dest: getattr(args, a.dest, None) for a in group._group_actions}
    enhance_kwargs["rng"] = rng

    files, rel_path, dir_proc = find_files(args.input)
    copy_stream = torch.cuda.Stream(device=device)

    def read(path):
        audio, fs = load_audio(path)
        return path, fs, audio

    def infos():
        for p in files:
            path, fs, audio = read(p)
            cache[path] = audio
            yield path, fs, audio.shape[0], audio.shape[1]

    cache = {}
    n_done = 0
    for group in group_files(infos(), max(1, args.batch_size)):
        fs = group[0][1]
        rows = torch.cat([cache.pop(path) for path, *_ in group], dim=0).pin_memory()
        with torch.cuda.stream(copy_stream):
            dev_rows = rows.to(device, non_blocking=True)
        torch.cuda.current_stream(device).wait_stream(copy_stream)
        with torch.no_grad():
            x = resample(dev_rows, fs, model.fs)
            enh = model.enhance(x, **enhance_kwargs)
            enh = resample(enh, model.fs, fs)
        host = torch.empty(enh.shape, dtype=enh.dtype).pin_memory()
        host.copy_(enh, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        r0 = 0
        for path, _, ch, _ in group:
            save_audio(output_path_for(path, rel_path, dir_proc, args.output), host[r0:r0 + ch], fs)
            r0 += ch
            n_done += 1
    return n_done


if __name__ == "__main__":
    main()
