#!/bin/bash
# round 2, call s: tile shape of the persistent kernel on mid-size documents
mkdir -p gpurun_out
for mib in 8 16 32 48; do
  KERNELS=persistent NWS=4,8,16,24 timeout 300 python tools/quickbench.py $mib 2>&1 | tail -4
done
KERNELS=split NWS=8,16 timeout 300 python tools/quickbench.py 48 2>&1 | tail -2
KERNELS=split NWS=8,16 timeout 300 python tools/quickbench.py 96 2>&1 | tail -2
