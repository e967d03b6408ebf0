#!/bin/bash
# round 2, call w: persistent flatten2 -- parity suite, then variants (resident CTAs per SM) via tools/quickbench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/w_pytest.log
for v in "" fl8 fl12; do
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
SJB200_FLATTEN=1 KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/w_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/w_ncu_list.log 2>&1
grep -E "flatten" gpurun_out/w_launches.csv | tail -2 | cut -d, -f5,12- 
