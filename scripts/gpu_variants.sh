#!/bin/bash
# usage: gpu_variants.sh <tag> "<variant>|<bench args>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  i=$((i+1))
  var="${spec%%|*}"; args="${spec#*|}"
  SJB200_LIB_VARIANT=$var timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 $args > gpurun_out/bench_${tag}_$i.log 2>&1
  echo "bench[variant=$var $args] rc=$?"
  tail -1 gpurun_out/bench_${tag}_$i.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('   value=%.1f GB/s ms=%.4f roofline=%.4f kernel_ms=%.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_ms']))
except Exception as e:
    print('   parse error',e)
"
done
