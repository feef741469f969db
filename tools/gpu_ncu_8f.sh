#!/bin/bash
# ncu --set full of the §8f kernels (one launch each): pack_layer_kernel at the C5 shape, phase / scale-round / samplers at C3 / C2
set -u
TAG=${1:-r04}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pack_layer_kernel|pack_init_kernel' -s 2 -c 2 -o "$OUT/pack" python tools/time_pack.py > "$OUT/pack_ncu.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'phase_ntt_kernel|sample_ternary_kernel|sample_cbd_kernel|sample_uniform_kernel' -s 8 -c 4 -o "$OUT/crypto" python tools/time_crypto.py > "$OUT/crypto_ncu.log" 2>&1
for f in pack crypto; do python tools/ncu_summary.py "$OUT/$f.ncu-rep" > "$OUT/${f}_summary.txt" 2>&1; done
head -60 "$OUT/pack_summary.txt"; grep -E "^==|dram__bytes|gpu__time_duration|dram_throughput" "$OUT/crypto_summary.txt"
