#!/bin/bash
# NTT-only GPU visit: parity tests of the transforms, then A/B timing of the warp kernel against the CTA-tile kernel.
# Usage: bash tools/gpu_ntt_ab.sh <tag>
set -u
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
echo "== pytest ntt"; (time timeout 600 python -m pytest tests/test_ntt_gpu.py -x -q) > "$OUT/pytest_ntt.log" 2>&1; tail -15 "$OUT/pytest_ntt.log"
echo "== time_ntt (warp kernel)"; timeout 300 python tools/time_ntt.py > "$OUT/time_ntt_warp.txt" 2>&1; cat "$OUT/time_ntt_warp.txt"
echo "== time_ntt (tile kernel)"; B2_NTT_TILE=1 timeout 300 python tools/time_ntt.py > "$OUT/time_ntt_tile.txt" 2>&1; cat "$OUT/time_ntt_tile.txt"
