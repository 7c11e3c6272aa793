#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary3.txt; }
: > gpurun_out/summary3.txt
run s3_parity_tc5 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "tcgen05 or edge or batch_invariance or host_buffer or enhanced"
run s3_bench_tc5 600 python bench.py --lstm-impl tcgen05 --steps 5 --warmup 3 --no-cpu-baseline
run s3_bench_tc5_fast 600 python bench.py --lstm-impl tcgen05 --fast-math --steps 5 --warmup 3 --no-cpu-baseline
run s3_mma_b64 150 python bench.py --lstm-impl mma --batch 64 --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary3.txt
tail -n 30 gpurun_out/s3_parity_tc5.log
for f in s3_bench_tc5 s3_bench_tc5_fast s3_mma_b64; do echo "--- $f"; tail -n 2 gpurun_out/$f.log | cut -c1-1800; done
