#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "tcgen05 or edge or batch_invariance or host_buffer" > gpurun_out/s7_parity.log 2>&1; echo "parity rc=$?"
tail -n 14 gpurun_out/s7_parity.log | cut -c1-300
: > gpurun_out/ab7.txt
for cfg in "1 0" "0 0" "1 1" "1 2" "1 3" "1 0"; do
  set -- $cfg
  r=$(FSN_TC5_PAIR=$1 FSN_TC5_DEBUG=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], round(d['roofline']['frac'],3))" 2>&1 | tail -1)
  echo "pair=$1 debug=$2 -> lstm_ms, step_ms, frac: $r" | tee -a gpurun_out/ab7.txt
done
