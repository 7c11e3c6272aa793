#!/bin/bash
# round 2, call K: S = 6 column split (12-CTA clusters), small-batch table, long-clip error diagnostic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -s -k "column_split or command_line or edge_shapes or plus_default_config_golden" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2k_pytest.log | tail -2
grep -E "^FAILED|^E  |column split" gpurun_out/r2k_pytest.log | cut -c1-250 | head -30
timeout 600 python scripts/time_fsn.py 2>&1 | grep -E "FullSubNet_Plus B=|fullsubnet.Model B=" | tee gpurun_out/r2k_small_batch.txt
timeout 900 python tests/diag_long_clip.py 2>&1 | tail -6 | tee gpurun_out/r2k_long_clip_diag.txt
