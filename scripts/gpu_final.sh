#!/bin/bash
# Round-end style pass: GPU tests, smoke, both bench arms, launch list + ncu full capture of the dominant kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2>gpurun_out/final_bench.err; echo "bench rc=$?"; tail -n 1 gpurun_out/final_bench.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5d -s 1 -c 1 -o gpurun_out/prof_tc5d_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
timeout 600 python scripts/time_fsn.py > gpurun_out/other_configs.txt 2>&1; echo "time_fsn rc=$?"; cat gpurun_out/other_configs.txt | tail -12
