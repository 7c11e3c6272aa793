#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5r -s 1 -c 1 -o gpurun_out/r2_prof_lstm_tc5r -f python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:gemm_f16 -s 1 -c 1 -o gpurun_out/r2_prof_gemm_f16 -f python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > gpurun_out/r2l_ncu2.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
