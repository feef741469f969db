#!/bin/bash
# developer visit: where the BumbleBee 1024^3 server time goes behind the adapter
set -u
TAG=${1:-r04}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
B=integration/_build
run() { local name=$1; shift; (time env "$@" timeout 600 $B/bench_matmul_b2 -D --bfv -m 1024 -n 1024 -r 1024 -st 20) > "$OUT/$name.log" 2>&1; echo "---- $name"; grep -E "Server-Matmul|  Matmul|Pack LWEs|b2_adapter|^real" "$OUT/$name.log"; }
run plain B2_ADAPTER=1 B2_ADAPTER_STATS=1
run launch_blocking B2_ADAPTER=1 B2_ADAPTER_STATS=1 CUDA_LAUNCH_BLOCKING=1
run no_tma B2_ADAPTER=1 B2_ADAPTER_STATS=1 B2_NTT5=0
