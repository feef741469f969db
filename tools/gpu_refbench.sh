#!/bin/bash
# The reference's OWN benchmarks (test/bench/*.cu, unmodified) on top of libb2rns through the adapter, next to the same
# binaries with the adapter off (= the reference's device kernels recompiled for sm_100): C5 BumbleBee matmul and the
# C3 / C4 / C2 HE-operation tables.  Usage (under gpurun): bash tools/gpu_refbench.sh <tag> [big]
set -u
TAG=${1:-r02}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
B=integration/_build
run() { # name, adapter(1/0), command...
  local name=$1 ad=$2; shift 2
  (time B2_ADAPTER=$ad B2_ADAPTER_STATS=1 timeout 900 "$@") > "$OUT/${name}_adapter$ad.log" 2>&1
  echo "---- $name adapter=$ad"; grep -vE "^\s*$" "$OUT/${name}_adapter$ad.log" | tail -${TAILN:-30}
}
# the first process on a freshly booted box is slow in BOTH arms (page cache, first large allocations): warm up before anything is compared
B2_ADAPTER=0 timeout 300 $B/bench_matmul_b2 -D --bfv -m 128 -n 128 -r 128 -st 20 > /dev/null 2>&1
Q3=60,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,60
for ad in 1 0; do
  TAILN=45 run matmul128 $ad $B/bench_matmul_b2 -D --bfv -m 128 -n 128 -r 128 -st 20
  TAILN=45 run matmul1024 $ad $B/bench_matmul_b2 -D --bfv -m 1024 -n 1024 -r 1024 -st 20
  TAILN=60 run ckks_c3 $ad $B/troybench_b2 -D --ckks -N 65536 -q $Q3 -R 20 -W 3 --no-bench-encode-simd --no-bench-encode-polynomial
  TAILN=60 run bfv_c2 $ad $B/troybench_b2 -D --bfv -N 16384 -q 60,40,40,60 -t 20 -B 256 -R 5 -W 2 --no-bench-encode-simd --no-bench-encode-polynomial
done
TAILN=45 run matmul128_host 1 $B/bench_matmul_b2 -H --bfv -m 128 -n 128 -r 128 -st 20
