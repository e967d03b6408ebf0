#!/bin/bash
# 128-byte-lane classify kernel against the 64-byte-lane one: mid sizes, CJK / dense / runs documents
for z in 1 0; do
  echo "SJB200_WIDE=$z"
  SJB200_WIDE=$z SIZES=96,128,256,512 KERNELS=stream timeout 300 python tools/sizesweep.py 2>&1 | tail -4
  for cfg in cjk dense runs; do
    SJB200_WIDE=$z timeout 300 python bench.py --config $cfg --steps 50 --warmup 10 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('  $cfg', d['value'], d['ms_per_step'])"
  done
done
