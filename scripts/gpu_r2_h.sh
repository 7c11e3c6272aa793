#!/bin/bash
# round 2, GPU call H: the hand-written input-projection GEMM (layer-wise path), config 5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -s -k "layerwise or config5 or three_layer or sb_output_activation or enhance_spectrum" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2h_pytest.log | tail -2
grep -E "^FAILED|^E  |layer-wise|config5" gpurun_out/r2h_pytest.log | cut -c1-250 | head -30
timeout 900 python bench.py --config 5 --steps 10 --warmup 5 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > gpurun_out/r2h_bench_c5.json 2> gpurun_out/r2h_bench_c5.err; echo "bench c5 rc=$?"; tail -c 300 gpurun_out/r2h_bench_c5.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench_c5.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print({"value": round(d["value"]), "ms": round(d["ms_per_step"], 3), "fwd_ms": round(d["forward_only"]["ms_per_step"], 3), "k_ms": round(r["kernel_ms"], 3), "frac": round(r["frac"], 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3)})
except Exception as e:
    print("no json", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2h_launches_c5.csv python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > /dev/null 2>&1; echo "ncu c5 rc=$?"
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2h_launches_c5.csv") if l.startswith('"'))]
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[1:]:
    if "fsn::" in r[ki] and ("gemm_f16" in r[ki] or "tc5r" in r[ki]):
        print(f"  {float(r[vi].replace(',', '')) / 1e3:10.1f} us  {r[ki][:90]}")
PY
