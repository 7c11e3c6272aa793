#!/bin/bash
# round 2, GPU call G: full GPU suite after the column-split fix, bench stability, small-batch table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1800 python -m pytest tests -m gpu -q -s --maxfail=6 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2g_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2g_pytest.log | cut -c1-300 | head -20
grep -E "column split|enhance|sb_act" gpurun_out/r2g_pytest.log | cut -c1-200 | head -30
timeout 600 python scripts/time_fsn.py 2>&1 | tail -22 | tee gpurun_out/r2g_small_batch.txt
for i in 1 2; do
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2g_bench_c2_$i.json 2> gpurun_out/r2g_bench_c2_$i.err; echo "bench c2 run $i rc=$?"; tail -c 300 gpurun_out/r2g_bench_c2_$i.err
  python - gpurun_out/r2g_bench_c2_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print({"value": round(d["value"]), "ms": round(d["ms_per_step"], 3), "host_ms": round(d.get("host_enqueue_ms_per_step", 0), 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3),
           "fwd_ms": round(d.get("forward_only", {}).get("ms_per_step", 0), 3), "cabi_ms": round(d.get("e2e_cabi", {}).get("ms_per_step", 0), 3),
           "k_ms": round(r.get("kernel_ms", 0), 3), "k_plain": round(r.get("kernel_ms_without_overlap", 0), 3), "frac": round(r.get("frac", 0), 3), "share": round(r.get("kernel_share_of_step", 0), 3)})
    print("overlap experiment", d.get("front_overlap_experiment"))
except Exception as e:
    print("no bench json", e)
PY
done
