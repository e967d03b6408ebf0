#!/bin/bash
# the GPU parity suite (optionally: -k expression as first argument)
mkdir -p gpurun_out
if [ -n "$1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_k.log
else
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_all.log
fi
