#!/bin/bash
# compute-sanitizer passes on the small-config suite (memcheck, synccheck, racecheck) + the tmem_alloc minimal reproducer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; export PYTHONDONTWRITEBYTECODE=1
K="plus_small_vs_oracle or fsn_small_golden or sb_output_activation or enhance_spectrum or submit_fullsubnet or fused_post"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "$K" > gpurun_out/r2_san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -n 4 gpurun_out/r2_san_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "plus_small_vs_oracle or sb_output_activation" > gpurun_out/r2_san_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -n 4 gpurun_out/r2_san_synccheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "plus_small_vs_oracle and tcgen05-64" > gpurun_out/r2_san_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "Race reported|hazard|RACECHECK|passed|failed" gpurun_out/r2_san_racecheck.log | cut -c1-220 | head -12
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "
import ctypes as C
p = C.CDLL('tests/libfsn_probe.so')
print('tmem_alloc reproducer returned', p.fsn_probe_tmem_alloc())
" > gpurun_out/r2_san_racecheck_repro.log 2>&1; echo "racecheck reproducer rc=$?"
grep -E "Race reported|hazard|RACECHECK|returned" gpurun_out/r2_san_racecheck_repro.log | cut -c1-220 | head -8
