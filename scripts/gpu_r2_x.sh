#!/bin/bash
# round 2, evidence call: full GPU suite, bench lines of the three configs + reference arm, launch list, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1800 python -m pytest tests -m gpu -q -s --maxfail=6 > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2x_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2x_pytest.log | cut -c1-300 | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2x_bench_c2.json 2> gpurun_out/r2x_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2x_bench_c2.err
timeout 900 python bench.py --config 5 --steps 10 --warmup 5 > gpurun_out/r2x_bench_c5.json 2> gpurun_out/r2x_bench_c5.err; echo "bench c5 rc=$?"; tail -c 300 gpurun_out/r2x_bench_c5.err
timeout 900 python bench.py --config 4 > gpurun_out/r2x_bench_c4.json 2> gpurun_out/r2x_bench_c4.err; echo "bench c4 rc=$?"; tail -c 300 gpurun_out/r2x_bench_c4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_ref_c2.json 2> gpurun_out/r2x_ref_c2.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 110 --csv --log-file gpurun_out/r2x_launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > /dev/null 2>&1; echo "ncu c2 rc=$?"
python - <<'PY'
import json
for f in ("c2", "c5", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/r2x_bench_{f}.json").read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, {"value": round(d["value"]), "ms": round(d["ms_per_step"], 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3), "fwd": round(d.get("forward_only", {}).get("ms_per_step", 0), 3),
                  "k_ms": round(r.get("kernel_ms", 0), 3), "frac": round(r.get("frac", 0), 3), "share": round(r.get("kernel_share_of_step", 0), 3), "lat": d.get("latency_ms", {}).get("p50")})
        if f == "c2":
            print("  cudnn", d.get("cudnn_baseline")); print("  cpu", {k: v for k, v in d.get("cpu_baseline", {}).items() if k in ("value", "cores", "kind", "concurrent")})
            print("  overlap", {k: v for k, v in (d.get("front_overlap_experiment") or {}).items() if k in ("ms_per_step", "lstm_kernel_ms")})
    except Exception as e:
        print(f, "no json", e)
d = json.loads(open("gpurun_out/r2x_ref_c2.json").read().strip().splitlines()[-1]); print("ref", d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
PY
timeout 600 python scripts/time_fsn.py 2>&1 | tee gpurun_out/r2x_small_batch.txt | tail -16
