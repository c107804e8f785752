"""Host logic of the inference CLI (open_universe_b200/bin/enhance.py; reference bin/enhance.py:83-192):
file discovery, output paths, length-bucketed grouping, WAV fallback I/O, refusal to run without CUDA."""
import numpy as np
import pytest
import torch

from open_universe_b200.bin import enhance as cli


def test_group_files_buckets_rows_by_length_and_rate():
    infos = [("a", 16000, 1, 100), ("b", 16000, 1, 100), ("c", 16000, 2, 100), ("d", 8000, 1, 100),
             ("e", 16000, 1, 50), ("f", 16000, 1, 100), ("g", 16000, 5, 100)]
    groups = list(cli.group_files(iter(infos), 3))
    names = [[i[0] for i in g] for g in groups]
    assert sorted(sum(names, [])) == list("abcdefg")             # every file exactly once
    for g in groups:
        assert len({(i[1], i[3]) for i in g}) == 1               # one (fs, length) per group
        assert sum(i[2] for i in g) <= 3 or len(g) == 1          # row cap, oversized file alone
    assert ["a", "b"] in names and ["g"] in names
    assert [[i[0] for i in g] for g in cli.group_files(iter(infos), 1)] == [[n] for n in "abcdefg"]


def test_find_files_and_output_paths(tmp_path):
    (tmp_path / "in" / "sub").mkdir(parents=True)
    for name in ("in/x.wav", "in/sub/y.flac", "in/notes.txt"):
        (tmp_path / name).write_bytes(b"")
    files, rel, dir_proc = cli.find_files(tmp_path / "in")
    assert dir_proc and [f.name for f in files] == ["y.flac", "x.wav"] or [f.name for f in files] == ["x.wav", "y.flac"]
    out = cli.output_path_for(tmp_path / "in" / "sub" / "y.flac", rel, True, tmp_path / "out")
    assert out == tmp_path / "out" / "sub" / "y.flac" and out.parent.is_dir()
    files, rel, dir_proc = cli.find_files(tmp_path / "in" / "x.wav")
    assert not dir_proc and files == [tmp_path / "in" / "x.wav"]
    assert cli.output_path_for(files[0], rel, False, tmp_path / "out") == tmp_path / "out" / "x.wav"
    assert cli.output_path_for(files[0], rel, False, tmp_path / "z.wav") == tmp_path / "z.wav"


def test_wav_roundtrip_and_pcm_scaling(tmp_path):
    from scipy.io import wavfile
    x = (0.3 * torch.randn(2, 321)).clamp(-1, 1)
    cli.save_audio(tmp_path / "f.wav", x, 16000)
    y, fs = cli.load_audio(tmp_path / "f.wav")
    assert fs == 16000 and torch.equal(x, y)
    wavfile.write(str(tmp_path / "i.wav"), 8000, np.array([0, 16384, -32768], dtype=np.int16))
    y, fs = cli.load_audio(tmp_path / "i.wav")
    assert fs == 8000 and torch.allclose(y, torch.tensor([[0.0, 0.5, -1.0]]))


def test_help_is_deferred_only_with_model():
    argv = ["in", "out", "--model", "m", "--help"]
    assert cli.handle_help(argv) and "--help" not in argv
    argv = ["in", "out", "-h"]
    assert not cli.handle_help(argv) and "-h" in argv


def test_no_cpu_inference(tmp_path):
    with pytest.raises(ValueError):
        cli.main([str(tmp_path), str(tmp_path), "--device", "cpu"])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            cli.main([str(tmp_path), str(tmp_path)])
