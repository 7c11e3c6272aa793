#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "fsn or config4 or streaming" > gpurun_out/s11.log 2>&1; echo "rc=$?"; grep -E "passed|failed|\[full|\[stream|\[config" gpurun_out/s11.log | cut -c1-200
timeout 300 python scripts/time_fsn.py 2>&1 | tail -5
timeout 300 python scripts/bench_stream.py 1 2>/dev/null | tee gpurun_out/stream_b1.json
