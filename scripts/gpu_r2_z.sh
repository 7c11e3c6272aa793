#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests -m gpu -q -s -k "chained or extreme" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2z_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2z_pytest.log | cut -c1-250 | head -10
