#!/bin/bash
# round 2, GPU call C: re-run of the fixed tests, bench stability (3 runs), sanitizers, ncu --set full of the sub-band LSTM kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -s -k "enhance or sb_output or reference_inferencer or submit or forward_host or causal" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/r2c_pytest.log | tail -2
grep -E "^FAILED|^E  " gpurun_out/r2c_pytest.log | cut -c1-300 | head -20
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print({"value": round(d["value"]), "ms": round(d["ms_per_step"], 3), "host_ms": round(d.get("host_enqueue_ms_per_step", 0), 3), "e2e_ms": round(d["e2e"]["ms_per_step"], 3),
           "fwd_ms": round(d.get("forward_only", {}).get("ms_per_step", 0), 3), "cabi_ms": round(d.get("e2e_cabi", {}).get("ms_per_step", 0), 3),
           "k_ms": round(r.get("kernel_ms", 0), 3), "k_plain": round(r.get("kernel_ms_without_overlap", 0), 3), "frac": round(r.get("frac", 0), 3),
           "share": round(r.get("kernel_share_of_step", 0), 3), "clocks": d.get("clocks")})
    print("timeline", d.get("timeline_ms", {}).get("last_steps_of_value"))
except Exception as e:
    print("no bench json", e)
PY
}
for i in 1 2 3; do
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2c_bench_c2_$i.json 2> gpurun_out/r2c_bench_c2_$i.err; echo "bench c2 run $i rc=$?"; tail -c 300 gpurun_out/r2c_bench_c2_$i.err; show gpurun_out/r2c_bench_c2_$i.json
done
bash scripts/gpu_sanitize.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc5d -s 2 -c 1 -o gpurun_out/r2_prof_lstm_tc5d -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cudnn-baseline > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
