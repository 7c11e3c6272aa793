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
  for i in 1 2 3; do
    timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2d_bench_c2_$i.json 2> gpurun_out/r2d_bench_c2_$i.err; echo "bench c2 run $i rc=$?"; tail -c 300 gpurun_out/r2d_bench_c2_$i.err
    python - gpurun_out/r2d_bench_c2_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print({"value": round(d["value"]), "ms": round(d["ms_per_step"], 3), "host_ms": round(d.get("host_enqueue_ms_per_step", 0), 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3),
           "fwd_ms": round(d.get("forward_only", {}).get("ms_per_step", 0), 3), "cabi_ms": round(d.get("e2e_cabi", {}).get("ms_per_step", 0), 3),
           "k_ms": round(r.get("kernel_ms", 0), 3), "k_plain": round(r.get("kernel_ms_without_overlap", 0), 3), "frac": round(r.get("frac", 0), 3), "share": round(r.get("kernel_share_of_step", 0), 3)})
    print("timeline", d.get("timeline_ms", {}).get("last_steps_of_value"))
except Exception as e:
    print("no bench json", e)
PY
  done
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5d -s 2 -c 1 -o gpurun_out/r2_prof_lstm_tc5d_b -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
