#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
: > gpurun_out/ab2.txt
for dbg in 0 1 2 3 0; do
  r=$(FSN_TC5_DEBUG=$dbg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'])")
  echo "debug=$dbg -> lstm_ms, step_ms: $r" | tee -a gpurun_out/ab2.txt
done
