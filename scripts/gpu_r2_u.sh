#!/bin/bash
# second convolution: residual two chunks ahead; dwconv 4 channels per thread: parity + step time + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 1200 python -m pytest tests -m gpu -q -s -k "extreme or config5 or plus or causal or enhance or edge or batch_130 or multiple_waves or fsn_small or fsn_default or distinct or streaming" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2u_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2u_pytest.log | cut -c1-250 | head -20
grep -E "extreme|stages:|tcgen05 stress|64 distinct|30 s offline|config5" gpurun_out/r2u_pytest.log | cut -c1-200 | head -20
for rep in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment 2>/dev/null | grep '^{"metric' > gpurun_out/r2u_bench$rep.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2u_bench$rep.json"))
print("bench$rep", "value ms", d["ms_per_step"], "fwd", d.get("forward_only",{}).get("ms_per_step"), "lstm", d["roofline"].get("kernel_ms"), "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "clk", d["clocks"]["sm_mhz"])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cudnn-baseline --no-overlap-experiment > gpurun_out/r2u_ncu_bench.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2u_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:70], float(r[vi].replace(',',''))) for r in rows[start+2:] if len(r)>vi]
idx=[i for i,(k,v) in enumerate(seq) if 'lstm_tc5d' in k]
if len(idx)>=2:
    a,b=idx[-2]+1, idx[-1]+1
    tot=0
    for k,v in seq[a:b]:
        print(f"{v/1e3:9.1f} us  {k}"); tot+=v
    print("total us", tot/1e3)
PY
