#!/bin/bash
# ncu --set full captures only (kept small: gpurun_out/ must stay under 64 MiB).  Usage: bash tools/gpu_ncu.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass_kernel' -s 8 -c 4 -f -o "$OUT/prof_ntt" \
    python tools/prof_ntt.py > "$OUT/prof_ntt.log" 2>&1
tail -2 "$OUT/prof_ntt.log"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ks_' -s 1 -c 1 -f -o "$OUT/prof_ksip" \
    python tools/prof_ks.py 2 > "$OUT/prof_ksip.log" 2>&1
tail -2 "$OUT/prof_ksip.log"
ls -la "$OUT"
