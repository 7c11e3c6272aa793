#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5 -s 1 -c 1 -o gpurun_out/prof_tc5_r01c python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc5 -s 3 -c 2 -o gpurun_out/prof_gemm_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
