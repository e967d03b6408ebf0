#!/bin/bash
# round 2, call u: balanced flatten kernel -- parity suite, then A/B against the lane-per-word kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/u_pytest.log
for f in 2 1; do
  SJB200_FLATTEN=$f timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/u_bench_f$f.log 2>&1
  echo "flatten=$f rc=$?"; tail -1 gpurun_out/u_bench_f$f.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('   value=%.1f GB/s ms=%.4f frac=%.4f'%(d['value'],d['ms_per_step'],d['roofline']['frac']))"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/u_ncu_list.log 2>&1
grep -E "flatten|classify|scan" gpurun_out/u_launches.csv | tail -8 | cut -d, -f5,12- 
