#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plus_small_vs_oracle or fsn_small_golden or streaming or fused_post" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -n 6 gpurun_out/san_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plus_small_vs_oracle and tcgen05-64" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -n 6 gpurun_out/san_racecheck.log | cut -c1-200
