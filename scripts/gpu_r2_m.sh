#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -s -k "column_split or internal_batch or submit_wait or command_line or enhance" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2m_pytest.log | tail -2
grep -E "^FAILED|^E  |column split" gpurun_out/r2m_pytest.log | cut -c1-250 | head -30
timeout 600 python scripts/time_fsn.py 2>&1 | grep -E "FullSubNet_Plus B= [1248]:|fullsubnet.Model B=1" | tee gpurun_out/r2m_small_batch.txt
