#!/bin/bash
# developer visit: persistent cluster / DSMEM transform — parity with every launch routed through it, then timings
mkdir -p gpurun_out
{
  echo "== parity, all N=2^16 launches through the cluster kernel"
  B2_NTT_CLUSTER_MAX=1000000 timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_ntt_gpu.py tests/test_keyswitch_gpu.py tests/test_host_mirror_gpu.py -x -q -m gpu 2>&1 | tail -5
  echo "== timings"
  timeout 300 python tools/time_cluster.py
  B2_NTT_CLUSTER_MAX=1000000 timeout 300 python tools/time_cluster.py
  B2_NTT_CLUSTER_MAX=1000000 B2_NTT_CLUSTERS=16 timeout 300 python tools/time_cluster.py
  B2_NTT_CLUSTER_MAX=1000000 B2_NTT_CLUSTERS=8 timeout 300 python tools/time_cluster.py
} > gpurun_out/$1_cluster.txt 2>&1
tail -40 gpurun_out/$1_cluster.txt
