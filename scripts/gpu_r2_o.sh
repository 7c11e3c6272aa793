#!/bin/bash
# same-box A/B of two builds of the library (ab/libbase.so = the committed kernel, in-tree = the candidate)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -x -k "tcgen05 or column_split or default or plus or stream or config4 or fb_num" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2o_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2o_pytest.log | cut -c1-250 | head -20
for rep in 1 2 3; do
  for which in base cand; do
    if [ $which = base ]; then export FSN_B200_LIB=$PWD/ab/libbase.so; else unset FSN_B200_LIB; fi
    timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment 2>/dev/null | grep '^{"metric' > gpurun_out/r2o_$which$rep.json
    python - <<PY
import json
d=json.load(open("gpurun_out/r2o_$which$rep.json"))
print("$which$rep", "value ms", d["ms_per_step"], "fwd", d.get("forward_only",{}).get("ms_per_step"), "lstm", d["roofline"].get("kernel_ms"), "frac", d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"])
PY
  done
done
