#!/bin/bash
# round 2, call f: windowed stream pipeline with one shared-memory carve-out for all its kernels
mkdir -p gpurun_out
for cfg in "0 4" "0 2" "64 2" "128 2" "256 2" "64 3" "128 3" "512 2"; do
  set -- $cfg
  SJB200_WINDOW_MIB=$1 SJB200_CLASSIFY_CTAS=$2 KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1 | sed "s/^/win=$1 ctas=$2 /"
done
