#!/bin/bash
# round 2, GPU call B: fused unfold + fused post-processing validation, A/B against the packed path, launch lists
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1800 python -m pytest tests -m gpu -q -s --maxfail=8 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2b_pytest.log | tail -3
grep -E "^FAILED|^E  " gpurun_out/r2b_pytest.log | cut -c1-300 | head -40
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print({"value": round(d["value"]), "ms": round(d["ms_per_step"], 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3), "fwd_ms": round(d.get("forward_only", {}).get("ms_per_step", 0), 3),
           "cabi_ms": round(d.get("e2e_cabi", {}).get("ms_per_step", 0), 3), "k_ms": round(r.get("kernel_ms", 0), 3), "k_plain": round(r.get("kernel_ms_without_overlap", 0), 3),
           "frac": round(r.get("frac", 0), 3), "share": round(r.get("kernel_share_of_step", 0), 3), "launches": d.get("gpu_launches"), "clocks": d.get("clocks")})
    print("timeline", d.get("timeline_ms", {}).get("last_steps_of_value"))
except Exception as e:
    print("no bench json", e)
PY
}
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench c2 (fused unfold) rc=$?"; tail -c 300 gpurun_out/r2b_bench_c2.err; show gpurun_out/r2b_bench_c2.json
FSN_NO_XFUSE=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2b_bench_c2_packed.json 2> gpurun_out/r2b_bench_c2_packed.err; echo "bench c2 (packed images) rc=$?"; show gpurun_out/r2b_bench_c2_packed.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2b_bench_c2_b.json 2>/dev/null; echo "bench c2 (fused unfold, repeat) rc=$?"; show gpurun_out/r2b_bench_c2_b.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/r2b_launches_c5.csv python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline > /dev/null 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/r2b_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline > /dev/null 2>&1; echo "ncu c2 rc=$?"
python - <<'PY'
import csv
for tag in ("c5", "c2"):
    try:
        rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r2b_launches_{tag}.csv") if l.startswith('"'))]
        h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        print(tag, "launches:", len(rows) - 1)
        for r in rows[1:75]:
            print(f"  {float(r[vi].replace(',', '')) / 1e3:10.1f} us  {r[ki][:90]}")
    except Exception as e:
        print(tag, "no launch list", e)
PY
