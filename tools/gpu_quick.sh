#!/bin/bash
# quick GPU visit: parity tests + timings (+ optional bench).  Usage: bash tools/gpu_quick.sh <tag> [bench]
set -u
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
(time timeout 1200 python -m pytest tests -m gpu -x -q) > "$OUT/pytest_gpu.log" 2>&1; tail -15 "$OUT/pytest_gpu.log"
timeout 600 python tools/time_ntt.py > "$OUT/time_ntt.txt" 2>&1; cat "$OUT/time_ntt.txt"
for alt in troy-nova_b200/alt_*.so; do [ -f "$alt" ] && { echo "== $alt"; B2RNS_LIB=$PWD/$alt timeout 600 python tools/time_ntt.py 2>&1 | tee "$OUT/time_ntt_$(basename $alt .so).txt"; }; done
B2_KS_TIMING=1 timeout 300 python tools/prof_ks.py 3 2>&1 | tail -4 | tee "$OUT/ks_timing.txt"
if [ "${2:-}" = "bench" ]; then timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; cat "$OUT/bench.json"; tail -3 "$OUT/bench.err"; fi
