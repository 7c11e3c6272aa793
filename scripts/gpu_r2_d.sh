#!/bin/bash
# round 2, GPU call D: column-split tests + small-batch timings (1 GPU part), then the N-rank checks when more GPUs are visible
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py 2>&1 | tail -4
  NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/r2d_bench_n$NG.json 2> gpurun_out/r2d_bench_n$NG.err; echo "bench N=$NG rc=$?"
  grep -E "NVLS|AllGather|all_gather|Connected|Channel 00" gpurun_out/r2d_bench_n$NG.err | head -8
  tail -c 400 gpurun_out/r2d_bench_n$NG.err
  tail -n 1 gpurun_out/r2d_bench_n$NG.json | cut -c1-2500
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2d_bench_n1.json 2>/dev/null; echo "bench N=1 rc=$?"
  tail -n 1 gpurun_out/r2d_bench_n1.json | cut -c1-400
else
  timeout 900 python -m pytest tests -m gpu -q -s -k "column_split or enhance or edge_shapes or plus_small_vs_oracle" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
  grep -E "passed|failed" gpurun_out/r2d_pytest.log | tail -2
  grep -E "^FAILED|^E  |column split" gpurun_out/r2d_pytest.log | cut -c1-300 | head -40
  timeout 600 python scripts/time_fsn.py 2>&1 | tail -22 | tee gpurun_out/r2d_small_batch.txt
fi
