#!/bin/bash
# usage: gpu_stream_ab.sh [variants...]   -- kernel-only throughput of the stream pipeline for each variant build
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" == "default" ]; then unset SJB200_LIB_VARIANT; else export SJB200_LIB_VARIANT=$v; fi
  KERNELS=${KERNELS:-stream} NWS=16 timeout 200 python tools/quickbench.py ${SIZE_MIB:-1024} 2>&1 | tail -${TAILN:-1}
done
