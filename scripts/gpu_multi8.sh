#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist rc=$?"; tail -n 2 gpurun_out/dist_check_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"
tail -n 1 gpurun_out/bench_n$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('value','n_gpus','ms_per_step','rtf')}, d['e2e']['ms_per_step'], d['pipeline']['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/bench_ref_n$N.log | cut -c1-300
