#!/bin/bash
# ncu --set full with source-level stalls of the depth-wise conv kernel and the TSSE gate kernel (final state)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'dwconv_tm|tsse_norm' --launch-skip 2 -c 3 -f -o gpurun_out/r2y_dw python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2y_dw.ncu-rep
