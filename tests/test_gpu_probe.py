"""GPU: tcgen05 / TMEM / bulk-copy instruction forms used by the persistent LSTM kernel, checked in isolation."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu


def test_tcgen05_probe(built_lib):
    import os
    probe = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfsn_probe.so"))   # test-only target, built by build()
    probe.fsn_probe_tcgen05.argtypes = [C.POINTER(C.c_float), C.c_int32]
    rep = (C.c_float * 48)()
    n = probe.fsn_probe_tcgen05(rep, 48)
    assert n == 43, n
    err_ss, err_ts, err_mix = list(rep)[:3]
    labels = [f"N{N}/{'TS' if ts else 'SS'}/acc{a}" for N in (64, 128, 192, 256) for ts in (1, 0) for a in (1, 2) if not (a == 2 and N > 128)]
    print(f"\nprobe: err ss={err_ss:.2e} ts={err_ts:.2e} mix={err_mix:.2e}")
    print("probe cycles per tcgen05.mma (M=128,K=16): " + "  ".join(f"{l}={c:.1f}" for l, c in zip(labels, list(rep)[3:15])))
    print(f"probe2: accumulator drain (16 warps, 128x128 fp32) = {rep[15]:.0f} cycles; M=64 N=128 MMA: TS {rep[16]:.1f}  SS {rep[17]:.1f} cycles")
    names = ["N128 lane0", "N128 elect", "N128 lane0+wait", "N128 elect+wait", "N128 lane0+commit", "N128 elect+commit",
             "N128 lane0+wait+commit", "N128 elect+wait+commit", "N128 lane0 walkA", "N128 elect+wait+commit walkA",
             "N64 lane0", "N64 elect", "N64 elect+wait+commit", "N256 elect", "N256 elect+wait+commit", "N256 elect+wait+commit walkA"]
    print("probe4 issue loop, cycles per tcgen05.mma: " + "; ".join(f"{k}={rep[18 + i]:.1f}" for i, k in enumerate(names)))
    print(f"probe5 CTA pair (cta_group::2, M=256 N=128): err SS={rep[34]:.2e} TS={rep[35]:.2e}; cycles/MMA SS={rep[36]:.1f} TS={rep[37]:.1f}")
    print(f"probe5 N=64 pair MMAs (TS): err={rep[38]:.2e}; cycles/MMA N64x4={rep[39]:.1f} N64x8(two acc)={rep[40]:.1f} "
          f"N128x4={rep[41]:.1f} N128x8(two acc)={rep[42]:.1f}")
    assert rep[34] < 1e-4 and rep[35] < 1e-4 and rep[38] < 1e-4
    # 64-term fp16 dot products of |x| <= 0.5 accumulated in fp32: exact to ~1e-6
    assert err_ss < 1e-4 and err_ts < 1e-4 and err_mix < 2e-4
