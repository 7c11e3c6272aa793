#!/bin/bash
# 2-GPU validation of the sharded path (NCCL all-gather of enhanced waveforms) + weak-scaling bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist rc=$?"
tail -n 1 gpurun_out/bench_n2.log | cut -c1-2500; echo; tail -n 1 gpurun_out/bench_n1.log | cut -c1-2500; echo; tail -n 5 gpurun_out/dist_check.log
