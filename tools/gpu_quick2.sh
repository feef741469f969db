#!/bin/bash
# Short GPU-box visit: parity tests, bench line, per-size timings, keyswitch stage timing.  Usage: bash tools/gpu_quick2.sh <tag>
set -u
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
echo "== pytest -m gpu"; (time timeout 900 python -m pytest tests -m gpu -x -q) > "$OUT/pytest_gpu.log" 2>&1; tail -5 "$OUT/pytest_gpu.log"
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline > "$OUT/bench.json" 2> "$OUT/bench.err"; tail -c 3500 "$OUT/bench.json"; tail -3 "$OUT/bench.err"
echo "== time_ntt"; timeout 300 python tools/time_ntt.py > "$OUT/time_ntt.txt" 2>&1; cat "$OUT/time_ntt.txt"
B2_KS_TIMING=1 timeout 300 python tools/prof_ks.py 3 2>&1 | tail -3 | tee "$OUT/ks_timing.txt"
