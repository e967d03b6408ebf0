#!/bin/bash
# parity tests under both kernels, then bench both
mkdir -p gpurun_out
for k in persist tile; do
  SJB200_KERNEL=$k timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$k.log 2>&1; echo "pytest[$k] rc=$?"; tail -8 gpurun_out/pytest_$k.log
done
for spec in "persist|--warps 8" "persist|--warps 4" "tile|--warps 8" "persist|--warps 8 --no-utf8"; do
  k="${spec%%|*}"; args="${spec#*|}"
  SJB200_KERNEL=$k timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 $args > gpurun_out/bench_ab.log 2>&1
  echo "bench[$k $args] rc=$?"
  tail -1 gpurun_out/bench_ab.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('   value=%.1f GB/s ms=%.4f roofline=%.4f kernel_ms=%.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_ms']))
except Exception as e:
    print('   parse error',e)
"
done
