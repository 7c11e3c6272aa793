#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_buffer" > gpurun_out/s9_host.log 2>&1; echo "host rc=$?"; tail -n 3 gpurun_out/s9_host.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s9_bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/s9_bench.log | cut -c1-3000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5p -s 1 -c 1 -o gpurun_out/prof_tc5p_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_p.log 2>&1; echo "ncu rc=$?"
