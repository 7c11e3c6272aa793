#!/bin/bash
# Validation + timing of the layer-wise tcgen05 path (FSN_TC5R=1): small-config parity first (short watchdog), then config #5.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "layerwise_tcgen05_vs_oracle" 2>&1 | tail -25 > gpurun_out/layerwise_small.log; echo "small rc=${PIPESTATUS[0]}"; tail -12 gpurun_out/layerwise_small.log
if grep -q "passed" gpurun_out/layerwise_small.log && ! grep -q "failed" gpurun_out/layerwise_small.log; then
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "layerwise_tcgen05_config5" 2>&1 | tail -8 | tee gpurun_out/layerwise_config5.log
  FSN_TC5R=1 timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/layerwise_timing.txt
import os, sys, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "fullsubnet-plus_b200")]
import bench
from fsnplus_b200.model import FullSubNet_Plus
dev = "cuda:0"
def timeit(fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    big = dict(bench.default_cfg(), num_freqs=513, sb_model_hidden_size=512, fb_model_hidden_size=512)
    l = FullSubNet_Plus(**big, num_layers=3).to(dev).eval()
    xl = torch.rand(32, 1, 513, 94, device=dev)
    ms = timeit(lambda: l(xl, xl - 0.5, xl - 0.3))
    print(f"config #5 (F=513, H=512, L=3, B=32, T=94), layer-wise tcgen05: {ms:.2f} ms  ({l.last_lstm_impl()}; {32 * 525.9 / ms:.0f} TFLOP/s algorithmic)")
PY
fi
