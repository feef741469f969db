#!/bin/bash
# ncu captures only, exported to CSV on the box so gpurun_out/ stays under 64 MiB.  Usage: bash tools/gpu_ncu.sh <tag>
set -u
TAG=${1:-r01}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ntt_persistent' -s 4 -c 2 -f -o "$OUT/prof_ntt" python tools/prof_ntt.py > "$OUT/prof_ntt.log" 2>&1
tail -2 "$OUT/prof_ntt.log"
ncu -i "$OUT/prof_ntt.ncu-rep" --page raw --csv > "$OUT/prof_ntt_raw.csv" 2>/dev/null
ncu -i "$OUT/prof_ntt.ncu-rep" --page source --csv > "$OUT/prof_ntt_source.csv" 2>/dev/null
ncu -i "$OUT/prof_ntt.ncu-rep" --page details > "$OUT/prof_ntt_details.txt" 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'ks_' -s 1 -c 1 -f -o "$OUT/prof_ksip" python tools/prof_ks.py 2 > "$OUT/prof_ksip.log" 2>&1
tail -2 "$OUT/prof_ksip.log"
ncu -i "$OUT/prof_ksip.ncu-rep" --page raw --csv > "$OUT/prof_ksip_raw.csv" 2>/dev/null
ncu -i "$OUT/prof_ksip.ncu-rep" --page details > "$OUT/prof_ksip_details.txt" 2>/dev/null
rm -f "$OUT/prof_ksip.ncu-rep"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ntt_|ks_' -s 5 --csv --log-file "$OUT/ks_launches.csv" python tools/prof_ks.py 2 > "$OUT/ks_launches.log" 2>&1
gzip -f "$OUT/prof_ntt_source.csv"
ls -la "$OUT"; du -sh gpurun_out
