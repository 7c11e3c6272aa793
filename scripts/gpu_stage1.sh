#!/bin/bash
# First GPU pass: every stage in its own process under `timeout`, logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run probe 300 python -m pytest tests/test_gpu_probe.py -m gpu -q -s
run parity_mma 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mma or fsn or three_layer or golden_from_reference"
run parity_tc5_small 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "tcgen05 and small"
run parity_tc5_default 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "tcgen05 and default"
run parity_rest 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "edge or batch_invariance or host_buffer or enhanced"
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_mma 600 python bench.py --lstm-impl mma --steps 3 --warmup 3
run bench_tc5 600 python bench.py --lstm-impl tcgen05 --steps 5 --warmup 3
cat gpurun_out/summary.txt
for f in probe parity_mma parity_tc5_small parity_tc5_default parity_rest smoke; do echo "--- $f"; tail -n 25 gpurun_out/$f.log; done
for f in bench_mma bench_tc5; do echo "--- $f"; tail -n 3 gpurun_out/$f.log | cut -c1-3000; done
