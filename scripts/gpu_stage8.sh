#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s8_all.log 2>&1; echo "all rc=$?"; tail -n 4 gpurun_out/s8_all.log
for fm in "" "--accurate-math"; do
  r=$(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $fm 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], round(d['roofline']['frac'],3), d['e2e']['ms_per_step'])" 2>&1 | tail -1)
  echo "fast='$fm' -> lstm_ms, step_ms, frac, e2e_ms: $r"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"
