#!/bin/bash
# parity subset through the stream pipeline, then variants via quickbench.  usage: gpu_r2_y2.sh <variant> ...
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "stream or both_classify or dense or flatten_units or 1gib or 64mib" > gpurun_out/y2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/y2_pytest.log
for v in "$@"; do
  if [ "$v" == "default" ]; then v=""; fi
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
