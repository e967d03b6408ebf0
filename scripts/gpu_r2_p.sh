#!/bin/bash
# round 2, call p: carries taken from the previous chunk inside a run; run length 2 / 4 / 8
mkdir -p gpurun_out
for v in "" t8 t2 ""; do
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2p.log
