#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "norm" > gpurun_out/s6_norms.log 2>&1; echo "norms rc=$?"
timeout 900 python -m pytest tests/ -m gpu -q > gpurun_out/s6_all.log 2>&1; echo "all rc=$?"
tail -n 30 gpurun_out/s6_norms.log | cut -c1-300; tail -n 5 gpurun_out/s6_all.log
