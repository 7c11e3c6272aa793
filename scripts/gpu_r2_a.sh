#!/bin/bash
# round 2, GPU call A: full GPU test-suite, bench lines of the three configs, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
nvidia-smi -L; nproc
timeout 1800 python -m pytest tests -m gpu -q -s --maxfail=6 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/r2a_pytest.log | tail -3
grep -E "^\[|cIRM|rel-L2|vs " gpurun_out/r2a_pytest.log | tail -60
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench c2 rc=$?"
tail -c 600 gpurun_out/r2a_bench_c2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2a_bench_c2.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "roofline", "e2e", "forward_only", "e2e_cabi", "timeline_ms", "clocks", "cudnn_baseline", "cpu_baseline") if k in d})
except Exception as e:
    print("no bench json", e)
PY
timeout 900 python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; echo "bench c5 rc=$?"
tail -c 400 gpurun_out/r2a_bench_c5.err; cut -c1-1500 gpurun_out/r2a_bench_c5.json
timeout 900 python bench.py --config 4 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; echo "bench c4 rc=$?"
tail -c 400 gpurun_out/r2a_bench_c4.err; cut -c1-1500 gpurun_out/r2a_bench_c4.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref_c2.json 2> gpurun_out/r2a_ref_c2.err; echo "ref rc=$?"
cut -c1-1200 gpurun_out/r2a_ref_c2.json
