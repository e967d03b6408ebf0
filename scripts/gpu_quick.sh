#!/bin/bash
# usage: gpu_quick.sh [--tests] "<env assignments>|<bench args>" ...
mkdir -p gpurun_out
if [ "$1" == "--tests" ]; then
  shift
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_quick.log
fi
i=0
for spec in "$@"; do
  i=$((i+1))
  envs="${spec%%|*}"; args="${spec#*|}"
  env $envs timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 $args > gpurun_out/bench_quick_$i.log 2>&1
  echo "bench[$envs $args] rc=$?"
  tail -1 gpurun_out/bench_quick_$i.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('   value=%.1f GB/s ms=%.4f roofline=%.4f kernel_ms=%.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_ms']))
except Exception as e:
    print('   parse error',e)
"
done
