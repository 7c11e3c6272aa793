#!/bin/bash
# L2 eviction hints on the dominant kernel's bulk copies: parity subset, bench line, one ncu capture for the DRAM traffic.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 200 python -m pytest tests -m gpu -q -x -k "plus_default_config_golden or batch_invariance or batch_130 or gru_default or plus_small_vs_oracle" 2>&1 | tail -3
timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench: lstm_ms', d['roofline']['kernel_ms'], 'step_ms', d['ms_per_step'], 'frac', d['roofline']['frac'])" | tee gpurun_out/hint_bench.txt
timeout 300 ncu --set full --clock-control none -k regex:lstm_tc5d -s 1 -c 1 -o gpurun_out/prof_tc5d_hint python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hint.log 2>&1; echo "ncu rc=$?"
