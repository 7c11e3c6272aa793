#!/bin/bash
# round 2, multi-GPU call: N-rank correctness check + the bench at N = all visible GPUs (NCCL_DEBUG output kept out of the JSON)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py 2>&1 | grep dist_check
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 20 --warmup 5 2> gpurun_out/r2n_bench_n$NG.err | grep '^{"metric' > gpurun_out/r2n_bench_n$NG.json; echo "bench N=$NG rc=${PIPESTATUS[0]}"
python - gpurun_out/r2n_bench_n$NG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", round(d["e2e"]["ms_per_step"], 3), "fwd", round(d["forward_only"]["ms_per_step"], 3), "k_ms", round(d["roofline"]["kernel_ms"], 3))
print(d["config"]["step"][:200])
PY
