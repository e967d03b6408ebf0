#!/bin/bash
# usage: gpu_profile_env.sh <tag> "<env assignments>" [bench args]
tag=$1; envs=$2; shift; shift
mkdir -p gpurun_out
CMD="env $envs python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 $*"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage1_ -s 2 -c 1 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "profile rc=$?"; tail -2 gpurun_out/ncu_full_$tag.log | cut -c1-200
