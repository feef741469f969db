#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-size timings, ncu launch list, ncu --set full of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag> [skip-full]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
echo "== pytest -m gpu" ; (time timeout 1500 python -m pytest tests -m gpu -x -q) > "$OUT/pytest_gpu.log" 2>&1; tail -5 "$OUT/pytest_gpu.log"
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -2 "$OUT/smoke.log"
echo "== bench" ; timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; tail -c 3000 "$OUT/bench.json"; tail -3 "$OUT/bench.err"
echo "== time_ntt" ; timeout 600 python tools/time_ntt.py --ref > "$OUT/time_ntt.txt" 2>&1; cat "$OUT/time_ntt.txt"
if [ "${2:-}" != "skip-ncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/bench_under_ncu.log" 2>&1
python tools/summarize_launches.py "$OUT/launches.csv" > "$OUT/launches_summary.txt" 2>&1; cat "$OUT/launches_summary.txt"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ntt_pass_kernel' -s 8 -c 4 -f -o "$OUT/prof_ntt" \
    python tools/prof_ntt.py > "$OUT/prof_ntt.log" 2>&1
tail -3 "$OUT/prof_ntt.log"
KSL=$(python tools/prof_ks.py 1 | sed -n 's/launches per keyswitch: //p'); echo "keyswitch launches: $KSL"
timeout 1200 ncu --set full --clock-control none -k regex:'ntt_pass_kernel|ks_' -s ${KSL:-21} -c ${KSL:-21} -f -o "$OUT/prof_ks" \
    python tools/prof_ks.py 2 > "$OUT/prof_ks.log" 2>&1
tail -3 "$OUT/prof_ks.log"
fi
ls -la "$OUT"
