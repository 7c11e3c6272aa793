#!/bin/bash
# ncu --set full of the two 1x1-convolution kernels in their final state (coalesced epilogues)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'gemm_tc5_kernel' --launch-skip 2 -c 2 -f -o gpurun_out/r2y_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > gpurun_out/r2y2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2y_gemm.ncu-rep
