#!/bin/bash
# parity suite, 1 GiB quick bench, mid-size sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/ab_pytest.log
KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
SIZES=16,32,64,96,128,256 KERNELS=persistent,split,stream timeout 600 python tools/sizesweep.py
