#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/s12_all.log 2>&1; echo "all rc=$?"; grep -E "passed|failed|fast_math" gpurun_out/s12_all.log | cut -c1-200
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],3), round(d['e2e']['ms_per_step'],3), round(d['pipeline']['ms_per_step'],3))"
done
