#!/bin/bash
# round 2, call v: full ncu capture of one kernel.  usage: gpu_r2_v.sh <kernel regex> <tag> [env...]
mkdir -p gpurun_out
re=$1; tag=$2; shift 2
env "$@" ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$tag.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$tag.log
