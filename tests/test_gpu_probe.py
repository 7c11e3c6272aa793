"""GPU: tcgen05 / TMEM / bulk-copy instruction forms used by the persistent LSTM kernel, checked in isolation."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_tcgen05_probe(built_lib):
    rep = (C.c_float * 16)()
    n = built_lib.fsn_probe_tcgen05(rep, 16)
    assert n == 7, built_lib.fsn_last_error()
    err_ss, err_ts, err_mix, c_ts64, c_ss64, c_ts256, c_ss256 = list(rep)[:7]
    print(f"\nprobe: err ss={err_ss:.2e} ts={err_ts:.2e} mix={err_mix:.2e}; cycles/MMA ts64={c_ts64:.1f} ss64={c_ss64:.1f} "
          f"ts256={c_ts256:.1f} ss256={c_ss256:.1f}")
    # 64-term fp16 dot products of |x| <= 0.5 accumulated in fp32: exact to ~1e-6
    assert err_ss < 1e-4 and err_ts < 1e-4 and err_mix < 2e-4
