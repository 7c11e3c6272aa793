#!/bin/bash
# 2-GPU pass with the current kernels: weak-scaling bench lines, sharded-harness check, and the command line under torchrun
# (files dealt to the ranks) compared byte for byte with a single-process run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=$PWD/fullsubnet-plus_b200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "dist rc=$?"
D=gpurun_out/cli_demo; rm -rf $D; python scripts/make_cli_demo.py $D > /dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 -m fsnplus_b200.tools.inference -C $D/inference.toml -M $D/ckpt.tar -I $D/noisy -O $D/out2 --batch_size 4 > gpurun_out/cli_n2.log 2>&1; echo "cli n2 rc=$?"
timeout 600 python -m fsnplus_b200.tools.inference -C $D/inference.toml -M $D/ckpt.tar -I $D/noisy -O $D/out1 --batch_size 64 > gpurun_out/cli_n1.log 2>&1; echo "cli n1 rc=$?"
python - <<'PY' | tee gpurun_out/cli_compare.txt
import os, numpy as np
from scipy.io import wavfile
a, b = "gpurun_out/cli_demo/out1/enhanced_0058", "gpurun_out/cli_demo/out2/enhanced_0058"
fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
assert fa == fb and len(fa) == 29, (len(fa), len(fb))
worst = 0
for f in fa:
    x, y = wavfile.read(os.path.join(a, f))[1].astype(int), wavfile.read(os.path.join(b, f))[1].astype(int)
    worst = max(worst, int(np.abs(x - y).max()))
print(f"CLI: 29 files from 2 ranks (batch 4) vs 1 process (batch 64): max |diff| = {worst} LSB (batch invariance of the forward)")
PY
grep "overall rtf" gpurun_out/cli_n1.log gpurun_out/cli_n2.log
rm -rf $D
tail -n 1 gpurun_out/bench_n2.log | cut -c1-700; echo; tail -n 1 gpurun_out/bench_n1.log | cut -c1-400; echo; tail -n 3 gpurun_out/dist_check.log
