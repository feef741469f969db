#!/bin/bash
# Round-2 GPU-box visit: parity tests (incl. the reference's own suites on libb2rns), smoke, bench line, ncu launch lists with DRAM
# bytes for the headline transform and one C3 keyswitch (-> profiles/ntt_traffic.json, profiles/ks_traffic.json).
# Usage (repo root, under gpurun): bash tools/gpu_visit2.sh <tag>
set -u
TAG=${1:-r02}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
echo "== pytest -m gpu"; (time timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=6) > "$OUT/pytest_gpu.log" 2>&1; tail -12 "$OUT/pytest_gpu.log"
echo "== smoke"; timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; tail -2 "$OUT/smoke.log"
echo "== bench"; timeout -s KILL 700 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; tail -3 "$OUT/bench.err"; head -c 600 "$OUT/bench.json"; echo
echo "== bench --impl reference"; timeout -s KILL 400 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; head -c 400 "$OUT/bench_reference.json"; echo
echo "== ncu launch lists"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ntt' -s 4 -c 2 --csv --log-file "$OUT/ntt_launches.csv" python tools/prof_ntt.py > "$OUT/ntt_launches.log" 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ntt_|ks_' -s 5 --csv --log-file "$OUT/ks_launches.csv" python tools/prof_ks.py 2 > "$OUT/ks_launches.log" 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > "$OUT/bench_under_ncu.log" 2>&1
python tools/summarize_launches.py "$OUT/launches.csv" > "$OUT/launches_summary.txt" 2>&1; head -30 "$OUT/launches_summary.txt"
timeout -s KILL 200 python tools/time_ops.py > "$OUT/time_ops.txt" 2>&1; cat "$OUT/time_ops.txt"
ls -la "$OUT"
