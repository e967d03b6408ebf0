#!/bin/bash
# usage: gpu_sweep.sh <tag> [--tests] -- "<bench args 1>" "<bench args 2>" ...
tag=$1; shift
mkdir -p gpurun_out
if [ "$1" == "--tests" ]; then
  shift
  timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
  timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_$tag.log
fi
[ "$1" == "--" ] && shift
i=0
for args in "$@"; do
  i=$((i+1))
  timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 $args > gpurun_out/bench_${tag}_$i.log 2>&1
  echo "bench[$args] rc=$?"
  tail -1 gpurun_out/bench_${tag}_$i.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    print('   value=%.1f GB/s ms=%.4f roofline=%.4f kernel_ms=%.4f e2e=%.1f clocks=%s'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel_ms'],d['e2e']['value'],d['clocks']))
except Exception as e:
    print('   parse error',e)
"
done
