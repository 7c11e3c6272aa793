#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary5.txt; }
: > gpurun_out/summary5.txt
run s5_parity 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_probe.py -m gpu -q -s
: > gpurun_out/ab5.txt
for cfg in "0 0" "1 0" "2 0" "3 0" "0 1" "0 0"; do
  set -- $cfg
  extra=""; [ "$2" = "1" ] && extra="--fast-math"
  r=$(FSN_TC5_DEBUG=$1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $extra 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'])")
  echo "debug=$1 fast=$2 -> lstm_ms, step_ms, frac: $r" | tee -a gpurun_out/ab5.txt
done
run s5_ncu_list 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary5.txt; tail -n 22 gpurun_out/s5_parity.log | cut -c1-400
