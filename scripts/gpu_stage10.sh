#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "streaming or fused_post" > gpurun_out/s10.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/s10.log | cut -c1-300
timeout 300 python scripts/bench_stream.py 1 > gpurun_out/stream_b1.json 2> gpurun_out/stream_b1.err; echo "stream rc=$?"; cat gpurun_out/stream_b1.json; tail -3 gpurun_out/stream_b1.err
timeout 300 python scripts/bench_stream.py 8 > gpurun_out/stream_b8.json 2>/dev/null; cat gpurun_out/stream_b8.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], d['pipeline']['ms_per_step'])"
