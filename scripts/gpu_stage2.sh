#!/bin/bash
# Second GPU pass: extended probe, mma kernel timing, ncu launch list + full capture of the tcgen05 kernel, CPU thread sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?" | tee -a gpurun_out/summary2.txt; }
: > gpurun_out/summary2.txt
run probe2 300 python -m pytest tests/test_gpu_probe.py -m gpu -q -s
run mma_b4 200 python bench.py --lstm-impl mma --batch 4 --steps 2 --warmup 3 --no-cpu-baseline
run mma_b16 300 python bench.py --lstm-impl mma --batch 16 --steps 2 --warmup 3 --no-cpu-baseline
run cpu_sweep 400 python - <<'PY'
import sys, time, os, torch
sys.path[:0] = [".", "fullsubnet-plus_b200"]
import bench
from fsnplus_b200.model import FullSubNet_Plus
from fsnplus_b200.synth import synth_clips
from fsnplus_b200 import inference as inf
cfg = bench.default_cfg(); torch.manual_seed(0)
m = FullSubNet_Plus(**cfg).eval(); st = m.state_dict()
X = inf.stft(synth_clips(2)); spec = (X.abs().unsqueeze(1), X.real.unsqueeze(1).contiguous(), X.imag.unsqueeze(1).contiguous())
for th in (128, 64, 32, 16, 8, 4):
    ref = bench.CpuReference(st, cfg, threads=th); ref.run(spec, 1)
    t = ref.run(spec, 3); print(th, "threads: s/clip", [round(x, 3) for x in t], flush=True)
PY
run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run ncu_full 1200 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5 -s 1 -c 1 -o gpurun_out/prof_tc5_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary2.txt
for f in probe2 mma_b4 mma_b16 cpu_sweep; do echo "--- $f"; tail -n 12 gpurun_out/$f.log | cut -c1-1500; done
tail -n 5 gpurun_out/ncu_list.log | cut -c1-400; tail -n 5 gpurun_out/ncu_full.log | cut -c1-400
ls -la gpurun_out | tail -20
