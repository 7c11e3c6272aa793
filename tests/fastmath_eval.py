"""Accuracy of the tanh.approx gate path vs the ex2/rcp path on long sequences (cell-state error accumulation)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fullsubnet-plus_b200")]
from oracle import fsn_oracle as O
from fsnplus_b200.model import FullSubNet_Plus, Model
dev = "cuda:0"
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
def build(cls, cfg, params, **kw):
    m = cls(**cfg, **kw); m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}); return m.to(dev).eval()
with torch.no_grad():
    # (1) causal FullSubNet, 12 s clip (T = 751), default sizes, LSTM weights x1 and x3
    for scale in (1.0, 3.0):
        cfg = O.default_fsn_config(); cfg["norm_type"] = "cumulative_laplace_norm"
        p = O.make_params_fsn(cfg, seed=21, lstm_scale=scale)
        mag = np.abs(O.stft(O.synth_clips(1, num_samples=192000, seed0=77)))[:, None].astype(np.float32)
        ref = O.fullsubnet_forward(p, cfg, mag)
        for fm in (False, True):
            out = build(Model, cfg, p, fast_math=fm)(t(mag)).cpu().numpy()
            print(f"FSN T={mag.shape[-1]} lstm_scale={scale} fast_math={fm}: rel-L2 {O.rel_l2(out, ref):.3e}", flush=True)
    # (2) FullSubNet+ default config, 9 s clip (T = 563)
    cfg = O.default_plus_config()
    X = O.stft(O.synth_clips(1, num_samples=144000, seed0=5))
    mag, real, imag = np.abs(X)[:, None].astype(np.float32), X.real[:, None].astype(np.float32), X.imag[:, None].astype(np.float32)
    for scale in (1.0, 3.0):
        p = O.make_params_plus(cfg, seed=0, lstm_scale=scale)
        ref = O.fullsubnet_plus_forward(p, cfg, mag, real, imag)
        for fm in (False, True):
            out = build(FullSubNet_Plus, cfg, p, fast_math=fm)(t(mag), t(real), t(imag)).cpu().numpy()
            print(f"PLUS T={mag.shape[-1]} lstm_scale={scale} fast_math={fm}: rel-L2 {O.rel_l2(out, ref):.3e}", flush=True)
