#!/bin/bash
mkdir -p gpurun_out
for v in "" ""; do
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "stream or corpus or length or misalign or fixtures or known or dense or maximum or speculation" > gpurun_out/pytest_r2q.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2q.log
