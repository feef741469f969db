#!/bin/bash
# GPU-box run of the reference's own test suite on top of libb2rns (integration/_build/troytest_b2), next to the same binary
# with the adapter switched off (B2_ADAPTER=0 = the reference's device branch) so that failures can be attributed.
# Usage (repo root, under gpurun): bash tools/gpu_adapter.sh <tag> [gtest filter]
set -u
TAG=${1:-r02}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
FILTER=${2:-*Device*:*device*:*Multithread*:*MultiThread*}
T=integration/_build/troytest_b2
echo "== troytest on libb2rns (adapter on)"
(time B2_ADAPTER_STATS=1 timeout 1500 $T --gtest_filter="$FILTER" --gtest_brief=1) > "$OUT/troytest_b2.log" 2>&1
grep -E "^\[  (PASSED|FAILED|SKIPPED) |tests ran|b2_adapter\]|^real" "$OUT/troytest_b2.log" | head -60
echo "== same binary, adapter off (reference device branch)"
(time B2_ADAPTER=0 B2_ADAPTER_STATS=1 timeout 1500 $T --gtest_filter="$FILTER" --gtest_brief=1) > "$OUT/troytest_ref.log" 2>&1
grep -E "^\[  (PASSED|FAILED|SKIPPED) |tests ran|b2_adapter\]|^real" "$OUT/troytest_ref.log" | head -60
