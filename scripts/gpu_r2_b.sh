#!/bin/bash
mkdir -p gpurun_out
for v in "" abl1 abl2 abl3 abl5; do
  SJB200_LIB_VARIANT=$v KERNELS=fused timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
KERNELS=fused timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage1_fused -s 3 -c 1 -o gpurun_out/prof_fused1 -f python tools/quickbench.py 1024 > gpurun_out/ncu_full_fused1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_fused1.log
