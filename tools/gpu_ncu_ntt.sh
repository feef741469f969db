#!/bin/bash
# ncu --set full of the NTT kernel (forward + inverse launch), exported on the box.  Usage: bash tools/gpu_ncu_ntt.sh <tag> [kernel regex]
set -u
TAG=${1:-n}; K=${2:-ntt_}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 2 -f -o "$OUT/prof_ntt" python tools/prof_ntt.py > "$OUT/prof_ntt.log" 2>&1
tail -2 "$OUT/prof_ntt.log"
python tools/ncu_summary.py "$OUT/prof_ntt.ncu-rep" > "$OUT/prof_ntt_summary.txt" 2>&1
ncu -i "$OUT/prof_ntt.ncu-rep" --page details > "$OUT/prof_ntt_details.txt" 2>/dev/null
ncu -i "$OUT/prof_ntt.ncu-rep" --page source --csv > "$OUT/prof_ntt_source.csv" 2>/dev/null
gzip -f "$OUT/prof_ntt_source.csv"; rm -f "$OUT/prof_ntt.ncu-rep"
cat "$OUT/prof_ntt_summary.txt" | head -70
