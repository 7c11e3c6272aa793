"""CPU: host logic of the reference-compatible command line (fsnplus_b200.tools.inference) -- file discovery, WAV I/O, the
int16 output convention of base_inferencer.py:151-152, length bucketing, config handling.  No compute calls."""
import os

import numpy as np
import pytest

from oracle import fsn_oracle as O

REF_TOML = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference_reference.toml")).read()


def test_reference_toml_builds_the_model(tmp_path):
    """The keys of the reference's config/inference.toml (restated above) construct the drop-in class unchanged."""
    from fsnplus_b200.tools import inference as T
    from fsnplus_b200 import model as M
    p = tmp_path / "inference.toml"
    p.write_text(REF_TOML)
    cfg = T.load_toml(p)
    assert cfg["model"]["args"]["sb_output_activate_function"] is False
    cls = getattr(M, T.MODEL_PATHS[cfg["model"]["path"]])
    net = cls(**cfg["model"]["args"])
    assert set(net.state_dict().keys()) == set(O.make_params_plus(O.default_plus_config(), seed=0).keys())
    assert T.INFERENCE_TYPES[cfg["inferencer"]["type"]] is True
    assert T.MODEL_PATHS["fullsubnet.model.fullsubnet.Model"] == "Model"


def test_wav_roundtrip_and_scaling(tmp_path):
    from scipy.io import wavfile
    from fsnplus_b200.tools import inference as T
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(4000) * 0.1).astype(np.float32)
    # output convention (base_inferencer.py:151-152): peak lands on 0.8 * 32767, truncation toward zero like np.int16()
    pcm = T.to_int16(y)
    assert pcm.dtype == np.int16 and np.abs(pcm).max() == int(0.8 * 32767)
    assert np.array_equal(pcm, np.int16(0.8 * 32767 * y / np.max(np.abs(y))))
    T.write_wav_int16(tmp_path / "a.wav", pcm, 16000)
    rate, back = wavfile.read(tmp_path / "a.wav")
    assert rate == 16000 and np.array_equal(back, pcm)
    # reader: int16 -> /32768 (librosa.load semantics), float32 passthrough, stereo averaged, other rates resampled
    assert np.array_equal(T.read_wav(tmp_path / "a.wav", 16000), pcm.astype(np.float32) / 32768.0)
    wavfile.write(tmp_path / "f.wav", 16000, y)
    assert np.array_equal(T.read_wav(tmp_path / "f.wav", 16000), y)
    wavfile.write(tmp_path / "s.wav", 16000, np.stack([y, -y * 0.5], 1))
    assert np.allclose(T.read_wav(tmp_path / "s.wav", 16000), 0.25 * y, atol=1e-7)
    t = np.arange(8000) / 8000.0
    wavfile.write(tmp_path / "r.wav", 8000, np.sin(2 * np.pi * 440 * t).astype(np.float32))
    up = T.read_wav(tmp_path / "r.wav", 16000)
    assert up.shape == (16000,) and up.dtype == np.float32
    ref = np.sin(2 * np.pi * 440 * np.arange(16000) / 16000.0)
    assert np.abs(up[200:-200] - ref[200:-200]).max() < 5e-3


def test_find_files_and_buckets(tmp_path):
    from fsnplus_b200.tools import inference as T
    (tmp_path / "a" / "sub").mkdir(parents=True)
    (tmp_path / "b").mkdir()
    for rel in ("a/z.wav", "a/sub/m.WAV", "a/notes.txt", "b/c.wav"):
        (tmp_path / rel).write_bytes(b"")
    got = [p.relative_to(tmp_path).as_posix() for p in T.find_files([tmp_path / "a", tmp_path / "b"])]
    assert got == ["a/sub/m.WAV", "a/z.wav", "b/c.wav"]                  # sorted within each directory, directories in order
    (tmp_path / "b" / "x.flac").write_bytes(b"")
    with pytest.raises(NotImplementedError):
        T.find_files([tmp_path / "b"])
    with pytest.raises(FileNotFoundError):
        T.find_files([tmp_path / "missing"])
    batches = T.bucket_by_length([10, 20, 10, 10, 20, 30, 10], batch_size=3)
    assert batches == [[0, 2, 3], [6], [1, 4], [5]]
    assert sorted(i for b in batches for i in b) == list(range(7))


def test_unknown_inferencer_type_and_model_path(tmp_path):
    from fsnplus_b200.tools import inference as T
    p = tmp_path / "c.toml"
    p.write_text(REF_TOML.replace("mag_complex_full_band_crm_mask", "sub_band_crm_mask"))
    with pytest.raises(NotImplementedError):
        T.run(T.load_toml(p), tmp_path / "none.tar", tmp_path, device="cpu")
    with pytest.raises(NotImplementedError):
        T.build_model({"path": "some.other.Model", "args": {}}, tmp_path / "none.tar", "cpu")


def test_run_host_logic_with_rank_sharding(tmp_path, monkeypatch):
    """Host logic of run(): length buckets, round-robin batches per rank (RANK / WORLD_SIZE as torchrun sets them), output
    directory + file names + int16 convention.  The model is a stand-in that predicts a constant compressed cIRM (2, 0), i.e. a
    positive real gain: the enhanced waveform is a scaled copy of the input, and the gain cancels in the peak normalisation, so
    the int16 output is predictable from the input file alone."""
    import torch
    from scipy.io import wavfile
    from fsnplus_b200.tools import inference as T
    rng = np.random.default_rng(3)
    noisy = tmp_path / "noisy"
    noisy.mkdir()
    lens = [4096, 4096, 6000, 4096, 6000, 4096, 4096]
    for i, n in enumerate(lens):
        wavfile.write(noisy / f"f{i}.wav", 16000, (rng.standard_normal(n) * 0.05).astype(np.float32))
    cfg = T.load_toml(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference_reference.toml"))
    cfg["dataset"]["args"]["dataset_dir_list"] = [str(noisy)]

    def constant_mask(mag, real, imag):                       # compressed cIRM (2.0, 0): decompresses to a positive real gain
        m = torch.zeros(mag.size(0), 2, mag.size(2), mag.size(3))
        m[:, 0] = 2.0
        return m

    seen = {}
    for rank in (0, 1):
        monkeypatch.setenv("RANK", str(rank)); monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("LOCAL_RANK", str(rank))
        seen[rank] = T.run(cfg, tmp_path / "unused.tar", tmp_path / "out", batch_size=2, device="cpu", log=lambda *a: None,
                           model_and_epoch=(constant_mask, 12))
    # batches: [0,1] [3,5] [6] [2,4] -> rank 0 takes batches 0 and 2, rank 1 batches 1 and 3
    assert sorted(seen[0]) == ["f0", "f1", "f6"] and sorted(seen[1]) == ["f2", "f3", "f4", "f5"]
    out_dir = tmp_path / "out" / "enhanced_0012"
    assert sorted(p.name for p in out_dir.iterdir()) == [f"f{i}.wav" for i in range(7)]
    for i, n in enumerate(lens):
        rate, pcm = wavfile.read(out_dir / f"f{i}.wav")
        x = wavfile.read(noisy / f"f{i}.wav")[1]
        assert rate == 16000 and pcm.dtype == np.int16 and pcm.shape == (n,)
        want = T.to_int16(x)                                   # a positive gain cancels in the peak normalisation
        assert np.abs(pcm.astype(int) - want.astype(int)).max() <= 2, i     # STFT -> iSTFT round trip in fp32


def test_wav_length_matches_read_wav_without_decoding(tmp_path):
    """Bucketing uses the header-only length (mmap); it must equal what read_wav returns, also after resampling."""
    from scipy.io import wavfile
    from fsnplus_b200.tools import inference as T
    rng = np.random.default_rng(0)
    for rate, n in ((16000, 12345), (44100, 30001), (8000, 7999), (48000, 48000)):
        p = tmp_path / f"x_{rate}.wav"
        wavfile.write(p, rate, (rng.standard_normal(n) * 3000).astype(np.int16))
        assert T.wav_length(p, 16000) == len(T.read_wav(p, 16000)), (rate, n)
