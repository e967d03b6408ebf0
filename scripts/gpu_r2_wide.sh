#!/bin/bash
# 128-byte-lane classify kernel: parity suite, then A/B against the 64-byte-lane kernel, then the launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/wide_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/wide_pytest.log
for z in 1 0; do
  echo "SJB200_WIDE=$z"
  SJB200_WIDE=$z KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
SJB200_WIDE=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 4 --csv --log-file gpurun_out/wide_launches.csv python tools/quickbench.py 1024 > gpurun_out/wide_ncu.log 2>&1
grep -E "classify|flatten|scan" gpurun_out/wide_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}'
