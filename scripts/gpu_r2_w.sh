#!/bin/bash
# chained (programmatic dependent) launches of the front-end kernels: parity with the knob on (default), same-box A/B of the step and of B = 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -s -k "extreme or config5 or plus or causal or enhance or edge or batch_130 or multiple_waves or fsn_small or fsn_default or distinct or streaming or submit or host" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2w_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2w_pytest.log | cut -c1-250 | head -20
for rep in 1 2; do
  for pdl in 0 1; do
    FSN_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment 2>/dev/null | grep '^{"metric' > gpurun_out/r2w_bench_pdl${pdl}_$rep.json
    python - <<PY
import json
d=json.load(open("gpurun_out/r2w_bench_pdl${pdl}_$rep.json"))
print("pdl=$pdl rep $rep", "value ms %.3f" % d["ms_per_step"], "fwd %.3f" % d["forward_only"]["ms_per_step"], "lstm %.3f" % d["roofline"]["kernel_ms"], "front %.3f" % (d["forward_only"]["ms_per_step"] - d["roofline"]["kernel_ms_without_overlap"]), "e2e", round(d["e2e"]["value"]), "clk", d["clocks"]["sm_mhz"])
PY
  done
done
for pdl in 0 1; do echo "FSN_PDL=$pdl"; FSN_PDL=$pdl timeout 600 python scripts/time_fsn.py 2>&1 | grep -E "FullSubNet_Plus B= [1248]:|fullsubnet.Model B=1:|config #5"; done | tee gpurun_out/r2w_small_batch.txt
