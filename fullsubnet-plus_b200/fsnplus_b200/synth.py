"""Synthetic 16 kHz noisy clips for benchmarks (SURVEY.md 8d recipe: harmonic 'speech' with a syllabic on/off
envelope plus white noise at SNR U[-5, 20] dB, level -25 dBFS +- 10; mirrors the mixing recipe of the reference's
fullsubnet/dataset/dataset_train.py:130-182).  Vectorised torch; deterministic per (seed, index)."""
import math

import torch


def synth_clips(n, num_samples=48000, sr=16000, seed=1000, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    u = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(*s, generator=g)
    t = torch.arange(num_samples, dtype=torch.float64) / sr
    f0 = u(100., 300., n, 1)
    sig = torch.zeros(n, num_samples, dtype=torch.float64)
    nh = torch.randint(3, 6, (n, 1), generator=g)
    for h in range(1, 6):
        amp = u(0.3, 1.0, n, 1) / h * (nh >= h)
        sig += amp * torch.sin(2 * math.pi * f0 * h * t + u(0., 2 * math.pi, n, 1))
    env = (torch.sin(2 * math.pi * u(3., 6., n, 1) * t + u(0., 2 * math.pi, n, 1)) > -0.2).double()
    k = torch.hann_window(321, periodic=False, dtype=torch.float64)
    env = torch.nn.functional.conv1d(env[:, None], (k / k.sum())[None, None], padding=160)[:, 0]
    sig = sig * env
    sig = sig * (0.1 / (sig.abs().amax(dim=1, keepdim=True) + 1e-9))
    snr = u(-5., 20., n, 1)
    noise = torch.randn(n, num_samples, generator=g, dtype=torch.float64)
    ps, pn = sig.pow(2).mean(1, keepdim=True) + 1e-12, noise.pow(2).mean(1, keepdim=True)
    y = sig + noise * torch.sqrt(ps / (pn * 10 ** (snr / 10)))
    level = u(-35., -15., n, 1)
    y = y * (10 ** (level / 20) / (y.pow(2).mean(1, keepdim=True).sqrt() + 1e-12))
    return y.clamp(-0.99, 0.99).float().to(device)
