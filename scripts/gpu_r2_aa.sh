#!/bin/bash
# variants of the run length + size sweep with finer sizes (kernel-choice thresholds)
for v in "" tc8 tc2; do
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
SIZES=8,16,24,32,48,64,96,128,192 KERNELS=persistent,split,stream timeout 600 python tools/sizesweep.py
