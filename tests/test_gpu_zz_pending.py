"""GPU tests written after round 1's GPU budget was spent: they have not run on a B200 yet, so they are marked
xfail(strict=False) -- an XPASS in the round-end log is the validation, an xfail names the first thing to fix next round -- and the
file name makes them the last ones collected.

  * fb_num_neighbors > 0 (the full-band outputs are unfolded like the sub-band window; reference fullsubnet_plus.py:167-179,
    fullsubnet.py:90-91) against committed reference outputs.  The packer and the utterance-mean kernels were written generic in the
    neighbour count from the start.
  * the layer-wise tcgen05 path (k_lstm_tc5r.cu) under fullsubnet.Model with a cumulative norm: the combination exercises the
    unswizzled store of the cumulative packer, which the validated layer-wise tests (FullSubNet+, offline norm) do not reach.
"""
import numpy as np
import pytest
import torch

from oracle import fsn_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="added after the GPU budget of round 1 was spent; not yet run on a B200")]
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _small(H):
    c = O.default_plus_config()
    c.update(num_freqs=33, sb_num_neighbors=3, sb_model_hidden_size=H)
    return c


def test_plus_fb_num_neighbors_golden(built_lib, golden):
    from fsnplus_b200.model import FullSubNet_Plus
    gi, g = golden("plus_small"), golden("plus_small_fbn1")
    cfg = dict(_small(32), fb_num_neighbors=1)
    m = FullSubNet_Plus(**cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_params_plus(cfg, seed=6).items()}, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m(_t(gi["mag"]), _t(gi["real"]), _t(gi["imag"]))
    err = O.rel_l2(out.cpu().numpy(), g["out"])
    print(f"\n[FullSubNet+ fb_num_neighbors=1] cIRM {err:.3e}")
    assert err < 1e-3


def test_fsn_fb_num_neighbors_golden(built_lib, golden):
    from fsnplus_b200.model import Model
    gi, g = golden("plus_small"), golden("fsn_small_fbn2")
    cfg = O.default_fsn_config()
    cfg.update(num_freqs=33, sb_num_neighbors=3, sb_model_hidden_size=32, fb_model_hidden_size=48, fb_num_neighbors=2)
    m = Model(**cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_params_fsn(cfg, seed=6).items()}, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m(_t(gi["mag"]))
    err = O.rel_l2(out.cpu().numpy(), g["out"])
    print(f"\n[fullsubnet.Model fb_num_neighbors=2] cIRM {err:.3e}")
    assert err < 1e-3


def test_layerwise_path_under_fullsubnet_model_with_cumulative_norm(built_lib):
    from fsnplus_b200.model import Model
    cfg = O.default_fsn_config()
    cfg.update(num_freqs=33, sb_num_neighbors=3, sb_model_hidden_size=64, fb_model_hidden_size=48, norm_type="cumulative_laplace_norm")
    params = O.make_params_fsn(cfg, seed=23, num_layers=3)
    rng = np.random.default_rng(4)
    mag = np.abs(rng.standard_normal((5, 1, 33, 24)) * 0.05 + 0.01).astype(np.float32)
    ref = O.fullsubnet_forward(params, cfg, mag, num_layers=3)
    m = Model(**cfg, num_layers=3)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        out = m(_t(mag))
    assert m.last_lstm_impl() == "tcgen05"
    err = O.rel_l2(out.cpu().numpy(), ref)
    print(f"\n[fullsubnet.Model, 3 x 64 sub-band, cumulative_laplace_norm, layer-wise tcgen05] cIRM {err:.3e}")
    assert err < 1e-3
