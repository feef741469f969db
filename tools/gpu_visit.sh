#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench line, per-size timings, keyswitch stage timing, ncu launch list,
# ncu --set full of the persistent NTT kernel (fwd + inv) and the key inner product, exported to CSV on the box.
# Usage (repo root, under gpurun): bash tools/gpu_visit.sh <tag> [noncu]
set -u
TAG=${1:-r01}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
echo "== pytest -m gpu"; (time timeout 1200 python -m pytest tests -m gpu -x -q) > "$OUT/pytest_gpu.log" 2>&1; tail -5 "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -2 "$OUT/smoke.log"
echo "== bench"; timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; tail -c 3500 "$OUT/bench.json"; tail -3 "$OUT/bench.err"
echo "== time_ntt"; timeout 600 python tools/time_ntt.py ${REF:-} > "$OUT/time_ntt.txt" 2>&1; cat "$OUT/time_ntt.txt"
B2_KS_TIMING=1 timeout 300 python tools/prof_ks.py 3 2>&1 | tail -4 | tee "$OUT/ks_timing.txt"
echo "== time_ops"; timeout 300 python tools/time_ops.py > "$OUT/time_ops.txt" 2>&1; cat "$OUT/time_ops.txt"
if [ "${2:-}" != "noncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1
python tools/summarize_launches.py "$OUT/launches.csv" > "$OUT/launches_summary.txt" 2>&1; cat "$OUT/launches_summary.txt"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_persistent -s 4 -c 2 -f -o "$OUT/prof_ntt" python tools/prof_ntt.py > "$OUT/prof_ntt.log" 2>&1
tail -2 "$OUT/prof_ntt.log"
ncu -i "$OUT/prof_ntt.ncu-rep" --page raw --csv > "$OUT/prof_ntt_raw.csv" 2>/dev/null
ncu -i "$OUT/prof_ntt.ncu-rep" --page source --csv > "$OUT/prof_ntt_source.csv" 2>/dev/null
ncu -i "$OUT/prof_ntt.ncu-rep" --page details > "$OUT/prof_ntt_details.txt" 2>/dev/null
gzip -f "$OUT/prof_ntt_source.csv"; rm -f "$OUT/prof_ntt.ncu-rep"
timeout 600 ncu --set full --clock-control none -k regex:ntt_cluster -s 12 -c 2 -f -o "$OUT/prof_cluster" python tools/prof_small.py > "$OUT/prof_cluster.log" 2>&1
ncu -i "$OUT/prof_cluster.ncu-rep" --page details > "$OUT/prof_cluster_details.txt" 2>/dev/null
python tools/ncu_summary.py "$OUT/prof_cluster.ncu-rep" > "$OUT/prof_cluster_summary.txt" 2>&1; rm -f "$OUT/prof_cluster.ncu-rep"
timeout 900 ncu --set full --clock-control none -k regex:'ks_' -s 1 -c 1 -f -o "$OUT/prof_ksip" python tools/prof_ks.py 2 > "$OUT/prof_ksip.log" 2>&1
tail -2 "$OUT/prof_ksip.log"
ncu -i "$OUT/prof_ksip.ncu-rep" --page raw --csv > "$OUT/prof_ksip_raw.csv" 2>/dev/null
ncu -i "$OUT/prof_ksip.ncu-rep" --page details > "$OUT/prof_ksip_details.txt" 2>/dev/null
rm -f "$OUT/prof_ksip.ncu-rep"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ntt_|ks_' -s 5 --csv --log-file "$OUT/ks_launches.csv" python tools/prof_ks.py 2 > "$OUT/ks_launches.log" 2>&1
fi
ls -la "$OUT"; du -sh gpurun_out
