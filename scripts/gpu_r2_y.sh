#!/bin/bash
# round 2, call y: parity suite + variants.  usage: gpu_r2_y.sh <variant> ...   ("default" = the product build)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/y_pytest.log
for v in "$@"; do
  if [ "$v" == "default" ]; then v=""; fi
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/y_ncu_list.log 2>&1
grep -E "flatten|classify|scan|persist" gpurun_out/y_launches.csv | tail -4 | cut -d, -f5,12- 
