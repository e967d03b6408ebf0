#!/bin/bash
# round 2, call g: upper bound of any overlapped organisation -- classify and flatten of the same document side by side, no dependency
mkdir -p gpurun_out
for cfg in "0 4 0" "0 2 1" "0 3 1" "0 4 1" "0 1 1"; do
  set -- $cfg
  SJB200_WINDOW_MIB=$1 SJB200_CLASSIFY_CTAS=$2 SJB200_EXPERIMENT_CORUN=$3 KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1 | sed "s/^/win=$1 ctas=$2 corun=$3 /"
done
