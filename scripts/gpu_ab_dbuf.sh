#!/bin/bash
# Same-box A/B of the double-buffered pair kernel (FSN_TC5_PAIR=2, default) against the single-accumulator pair kernel (=1):
# parity tests of the tcgen05 path first, then the bench line of each.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 420 python -m pytest tests -m gpu -q -x -s -k "plus_default_config_golden or plus_small_vs_oracle or batch_invariance or batch_130 or large_batch or gru_plus_small or gru_default or edge_shapes or accurate_gate" 2>&1 | tail -40 > gpurun_out/ab_dbuf_tests.log
tail -5 gpurun_out/ab_dbuf_tests.log
: > gpurun_out/ab_dbuf.txt
for pair in 2 1 2 1; do
  r=$(FSN_TC5_PAIR=$pair timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])")
  echo "FSN_TC5_PAIR=$pair -> lstm_ms, step_ms, frac, clocks: $r" | tee -a gpurun_out/ab_dbuf.txt
done
