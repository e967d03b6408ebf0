#!/bin/bash
# full ncu capture of one kernel of the stream pipeline via quickbench.  usage: gpu_r2_v2.sh <kernel regex> <tag> [env...]
mkdir -p gpurun_out
re=$1; tag=$2; shift 2
env KERNELS=stream "$@" ncu --set full --clock-control none --import-source on -k regex:$re -s 2 -c 1 -f -o gpurun_out/prof_$tag python tools/quickbench.py 1024 > gpurun_out/ncu_$tag.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_$tag.log
