#!/bin/bash
# round 2, call z: all bench configurations + size sweep with the current build
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 2 "$@" > gpurun_out/z_$tag.log 2>&1; echo "$tag rc=$?"; tail -1 gpurun_out/z_$tag.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('   value=%.1f %s ms=%.4f frac=%s e2e=%s'%(d['value'],d['unit'],d['ms_per_step'],d.get('roofline',{}).get('frac'),d.get('e2e',{}).get('value')))
except Exception as e: print('   parse error', e)"; }
run doc1g
run noutf8 --no-utf8
run cjk --config cjk
run dense --config dense
run runs --config runs
run small --config small
KERNELS=auto,persistent,split,stream timeout 600 python tools/sizesweep.py > gpurun_out/z_sizesweep.log 2>&1; cat gpurun_out/z_sizesweep.log
