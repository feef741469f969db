#!/bin/bash
# developer visit: the packing / crypto seams behind the adapter — the reference's matmul / conv2d / LWE / Encryptor suites, then the reference's
# own bench_matmul (C5) with the adapter on / off
set -u
TAG=${1:-r04}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
B=integration/_build
timeout 900 python -m pytest tests/test_adapter_gpu.py -q -m gpu -k "matmul or lwe or encryptor" > "$OUT/adapter_pytest.log" 2>&1; tail -5 "$OUT/adapter_pytest.log"
for ad in 1 0; do
  for sz in 128 1024; do
    (time B2_ADAPTER=$ad B2_ADAPTER_STATS=1 timeout 600 $B/bench_matmul_b2 -D --bfv -m $sz -n $sz -r $sz -st 20) > "$OUT/matmul${sz}_adapter$ad.log" 2>&1
    echo "---- matmul $sz adapter=$ad"; grep -vE "^\s*$" "$OUT/matmul${sz}_adapter$ad.log" | tail -32
  done
done
