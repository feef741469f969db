#!/bin/bash
set -u
TAG=${1:-p}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ntt_persistent' -s 4 -c 2 -f -o "$OUT/prof_ntt" python tools/prof_ntt.py > "$OUT/prof_ntt.log" 2>&1
tail -2 "$OUT/prof_ntt.log"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ntt_|ks_' -s 6 --csv --log-file "$OUT/ks_launches.csv" python tools/prof_ks.py 2 > "$OUT/ks_launches.log" 2>&1
grep -E "ntt_|ks_" "$OUT/ks_launches.csv" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | sed 's/(b2::NttJob.*)//' | cut -c1-220
