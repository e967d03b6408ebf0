#!/bin/bash
# round 2, call x: variants of the stream pipeline via tools/quickbench.py.  usage: gpu_r2_x.sh <variant> ...
for v in "$@"; do
  if [ "$v" == "default" ]; then v=""; fi
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
