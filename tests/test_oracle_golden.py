"""CPU: the numpy oracle must reproduce every golden vector generated from the unmodified reference
(tests/golden/make_golden.py).  Tolerance: float64 restatement vs float64 reference stored as float32
-> 1e-6 relative L2 (storage rounding only)."""
import numpy as np
import pytest

from oracle import fsn_oracle as O

TOL = 1e-6


def small_plus_cfg():
    c = O.default_plus_config()
    c.update(num_freqs=33, sb_num_neighbors=3, sb_model_hidden_size=32)
    return c


def small_fsn_cfg(norm):
    c = O.default_fsn_config()
    c.update(num_freqs=33, sb_num_neighbors=3, sb_model_hidden_size=32, fb_model_hidden_size=48, norm_type=norm)
    return c


def test_plus_small_all_stages(golden):
    g = golden("plus_small")
    cfg = small_plus_cfg()
    st = {}
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=3), cfg, g["mag"], g["real"], g["imag"], stages=st)
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_in"], g["fb_in"]) < TOL
    assert O.rel_l2(st["fb_out"], g["fb_out"]) < TOL
    assert O.rel_l2(st["sb_in"], g["sb_in"]) < TOL


@pytest.mark.parametrize("attn", ["SE", "ECA", "CBAM"])
def test_plus_small_other_attentions(golden, attn):
    """channel_attention_model variants of fullsubnet_plus.py:52-70 (SE is the reference constructor's default)."""
    g, gi = golden(f"plus_small_{attn}"), golden("plus_small")
    cfg = dict(small_plus_cfg(), channel_attention_model=attn)
    st = {}
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=5), cfg, gi["mag"], gi["real"], gi["imag"], stages=st)
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_in"], g["fb_in"]) < TOL


@pytest.mark.parametrize("sub", [2, 3])
def test_plus_small_subband_num(golden, sub):
    """subband_num > 1 (fullsubnet_plus.py:146-153) -- runs in the reference only with ECA; F = 33 pads 1 bin (sub = 2) or a
    whole extra group (sub = 3)."""
    g, gi = golden(f"plus_small_ECA_sub{sub}"), golden("plus_small")
    cfg = dict(small_plus_cfg(), channel_attention_model="ECA", subband_num=sub)
    st = {}
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=5), cfg, gi["mag"], gi["real"], gi["imag"], stages=st)
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_in"], g["fb_in"]) < TOL


def test_fb_num_neighbors(golden):
    """fb_num_neighbors > 0: the full-band outputs are unfolded as well (fullsubnet_plus.py:167-179, fullsubnet.py:90-91)."""
    gi = golden("plus_small")
    cfg = dict(small_plus_cfg(), fb_num_neighbors=1)
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=6), cfg, gi["mag"], gi["real"], gi["imag"])
    assert O.rel_l2(out, golden("plus_small_fbn1")["out"]) < TOL
    cfg = dict(small_fsn_cfg("offline_laplace_norm"), fb_num_neighbors=2)
    out = O.fullsubnet_forward(O.make_params_fsn(cfg, seed=6), cfg, gi["mag"])
    assert O.rel_l2(out, golden("fsn_small_fbn2")["out"]) < TOL


def test_gru_variants(golden):
    """sequence_model = "GRU" (sequence_model.py:39-46) for both model classes, gates pushed towards saturation (x2)."""
    gi = golden("plus_small")
    cfg = dict(small_plus_cfg(), sequence_model="GRU")
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=8, lstm_scale=2.0), cfg, gi["mag"], gi["real"], gi["imag"])
    assert O.rel_l2(out, golden("plus_small_GRU")["out"]) < TOL
    cfg = dict(small_fsn_cfg("offline_laplace_norm"), sequence_model="GRU")
    st = {}
    out = O.fullsubnet_forward(O.make_params_fsn(cfg, seed=8, lstm_scale=2.0), cfg, gi["mag"], stages=st)
    g = golden("fsn_small_GRU")
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_out"], g["fb_out"]) < TOL


@pytest.mark.parametrize("norm", ["offline_laplace_norm", "cumulative_laplace_norm", "offline_gaussian_norm", "cumulative_layer_norm"])
def test_fsn_small_norms(golden, norm):
    g = golden(f"fsn_small_{norm}")
    cfg = small_fsn_cfg(norm)
    st = {}
    out = O.fullsubnet_forward(O.make_params_fsn(cfg, seed=4), cfg, g["mag"], stages=st)
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_out"], g["fb_out"]) < TOL


def test_lstm3(golden):
    g = golden("lstm3_small")
    p3 = O._lstm_params(np.random.default_rng(12), "sb_model", 10, 16, 3, 2)
    assert O.rel_l2(O.seq_lstm(g["x"].astype(np.float64), p3, "sb_model", 3, False), g["out"]) < TOL


def test_plus_default_and_pipeline(golden):
    g = golden("plus_default")
    cfg = O.default_plus_config()
    clips = O.synth_clips(1)
    X = O.stft(clips)
    # numpy STFT == torch.stft of the reference (inputs stored as float32)
    assert O.rel_l2(np.abs(X)[:, None], g["mag"]) < 1e-6
    assert O.rel_l2(X.real[:, None], g["real"]) < 1e-5
    st = {}
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=0), cfg, g["mag"], g["real"], g["imag"], stages=st)
    assert out.shape == (1, 2, 257, 188)
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_in"], g["fb_in"]) < TOL
    assert O.rel_l2(st["fb_out"], g["fb_out"]) < TOL
    # decompress + complex multiply + iSTFT (inferencer.py:152-158)
    enh = O.istft(O.enhance(X, g["out"].astype(np.float64)), length=clips.shape[1])
    assert O.rel_l2(enh, g["enhanced"]) < 1e-5


def test_plus_default_stress(golden):
    g, gi = golden("plus_default_stress"), golden("plus_default")
    cfg = O.default_plus_config()
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=0, lstm_scale=3.0), cfg, gi["mag"], gi["real"], gi["imag"])
    assert O.rel_l2(out, g["out"]) < TOL


def test_fsn_default(golden):
    g, gi = golden("fsn_default"), golden("plus_default")
    cfg = O.default_fsn_config()
    st = {}
    out = O.fullsubnet_forward(O.make_params_fsn(cfg, seed=1), cfg, gi["mag"], stages=st)
    assert O.rel_l2(out, g["out"]) < TOL
    assert O.rel_l2(st["fb_out"], g["fb_out"]) < TOL


def test_sb_mean_identity():
    """K4 identity: the mean of the unfolded tensor from row sums and reflection counts."""
    rng = np.random.default_rng(0)
    F, T, N = 37, 11, 5
    x = rng.standard_normal((2, 1, F, T))
    direct = O.unfold(x, N).sum(axis=(1, 2, 3, 4))
    S = x[:, 0].sum(axis=2)
    idx = np.arange(F)[:, None] + np.arange(2 * N + 1)[None, :] - N
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx > F - 1, 2 * (F - 1) - idx, idx)
    assert np.allclose(S[:, idx].sum(axis=(1, 2)), direct, rtol=1e-12)


def test_flops_match_survey():
    f = O.flops_plus(O.default_plus_config(), 188)
    assert abs(f["total"] / 1e9 - 180.477) < 0.01
    assert abs(f["subband"] / 1e9 - 177.982) < 0.01
    g = O.flops_fsn(O.default_fsn_config(), 188)
    assert abs(g["total"] / 1e9 - 179.127) < 0.01


def test_torch_port_matches_reference(golden):
    """The torch-CPU port that bench.py times as the CPU baseline reproduces the reference (fp32 noise floor)."""
    import torch
    from oracle.torch_port import TorchPort
    g = golden("plus_default")
    cfg = O.default_plus_config()
    port = TorchPort(O.make_params_plus(cfg, seed=0), cfg, "plus")
    t = lambda k: torch.from_numpy(g[k])
    out = port.forward(t("mag"), t("real"), t("imag")).numpy()
    assert O.rel_l2(out, g["out"]) < 1e-4
    gf = golden("fsn_default")
    fcfg = O.default_fsn_config()
    out = TorchPort(O.make_params_fsn(fcfg, seed=1), fcfg, "fsn").forward(t("mag")).numpy()
    assert O.rel_l2(out, gf["out"]) < 1e-4


@pytest.mark.parametrize("norm", ["offline_laplace_norm", "cumulative_laplace_norm", "offline_gaussian_norm", "cumulative_layer_norm"])
def test_torch_port_fp64_all_norms_pinned(golden, norm):
    """The torch port in float64 is the truth of the long-clip GPU parity tests (the numpy oracle needs minutes for a 30 s clip):
    pin it to the reference goldens for every norm_type at the fp64 noise floor."""
    import torch
    from oracle.torch_port import TorchPort
    g = golden(f"fsn_small_{norm}")
    c = O.default_fsn_config()
    c.update(num_freqs=33, sb_num_neighbors=3, sb_model_hidden_size=32, fb_model_hidden_size=48, norm_type=norm)
    port = TorchPort(O.make_params_fsn(c, seed=4), c, "fsn", dtype=torch.float64)
    out = port.forward(torch.from_numpy(g["mag"])).numpy()
    assert O.rel_l2(out, g["out"]) < 1e-6          # goldens are stored in float32


def test_plus_small_causal_tcn(golden):
    """Causal FullSubNet+ variant: TCNBlock(causal=True) (causal_conv.py:74-75,104-105) in the three full-band
    models; golden from the reference's own block class (tests/golden/make_golden.py: causal_plus)."""
    g, gi = golden("plus_small_causal"), golden("plus_small")
    cfg = dict(small_plus_cfg(), causal_tcn=True)
    st = {}
    out = O.fullsubnet_plus_forward(O.make_params_plus(cfg, seed=14), cfg, gi["mag"], gi["real"], gi["imag"], stages=st)
    assert O.rel_l2(out, g["out"]) < TOL and O.rel_l2(st["fb_out"], g["fb_out"]) < TOL
    import torch
    from oracle.torch_port import TorchPort
    port = TorchPort(O.make_params_plus(cfg, seed=14), cfg, "plus", dtype=torch.float64)
    outs = np.concatenate([port.forward(*(torch.from_numpy(gi[k][b:b + 1]) for k in ("mag", "real", "imag"))).numpy() for b in range(3)])
    assert O.rel_l2(outs, g["out"]) < TOL
