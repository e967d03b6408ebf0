#!/bin/bash
# round 2, call o: flatten kernel with 1 / 2 / 4 chunks per warp; parity of the default
mkdir -p gpurun_out
for v in cpw1 "" cpw4 cpw1 ""; do
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
SJB200_LIB_VARIANT= KERNELS=split,stream timeout 300 python tools/quickbench.py 64 2>&1 | tail -2
SJB200_LIB_VARIANT=cpw1 KERNELS=split,stream timeout 300 python tools/quickbench.py 64 2>&1 | tail -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_r2o.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2o.log
