#!/bin/bash
# A/B timing of tcgen05 LSTM tuning knobs on one box (same process conditions, alternating order).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
: > gpurun_out/ab.txt
for rep in 1 2; do
 for cfg in "0 0" "1 0" "0 4" "0 3"; do
  set -- $cfg
  r=$(FSN_TC5_ELECT=$1 FSN_TC5_NSTAGE=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['clocks'])")
  echo "elect=$1 nstage_cap=$2 rep=$rep -> lstm_ms, step_ms, clocks: $r" | tee -a gpurun_out/ab.txt
 done
done
