#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary4.txt; }
: > gpurun_out/summary4.txt
run s4_parity_small 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "small"
run s4_parity_all 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s
run s4_bench 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run s4_ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary4.txt
tail -n 40 gpurun_out/s4_parity_small.log | cut -c1-300
tail -n 40 gpurun_out/s4_parity_all.log | cut -c1-300
tail -n 2 gpurun_out/s4_bench.log | cut -c1-1500
