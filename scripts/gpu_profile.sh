#!/bin/bash
# ncu launch list + full capture of the stage-1 kernel.  usage: gpu_profile.sh <tag> [bench args...]
tag=$1; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1 $*"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage1_ -s 2 -c 2 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
echo "profile rc=$?"; tail -2 gpurun_out/plain_$tag.log | cut -c1-300; tail -5 gpurun_out/ncu_full_$tag.log
