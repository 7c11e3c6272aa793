"""GPU: tcgen05 / TMEM / bulk-copy instruction forms used by the persistent LSTM kernel, checked in isolation."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_tcgen05_probe(built_lib):
    rep = (C.c_float * 32)()
    n = built_lib.fsn_probe_tcgen05(rep, 32)
    assert n == 15, built_lib.fsn_last_error()
    err_ss, err_ts, err_mix = list(rep)[:3]
    labels = [f"N{N}/{'TS' if ts else 'SS'}/acc{a}" for N in (64, 128, 192, 256) for ts in (1, 0) for a in (1, 2) if not (a == 2 and N > 128)]
    print(f"\nprobe: err ss={err_ss:.2e} ts={err_ts:.2e} mix={err_mix:.2e}")
    print("probe cycles per tcgen05.mma (M=128,K=16): " + "  ".join(f"{l}={c:.1f}" for l, c in zip(labels, list(rep)[3:15])))
    # 64-term fp16 dot products of |x| <= 0.5 accumulated in fp32: exact to ~1e-6
    assert err_ss < 1e-4 and err_ts < 1e-4 and err_mix < 2e-4
