#!/bin/bash
# round 2, call i: ncu of the current build -- launch list of the bench command and one full capture (with source) of a document pass
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_r2i.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2i.csv $CMD > gpurun_out/ncu_list_r2i.log 2>&1; echo "list rc=$?"
KERNELS=stream timeout 600 python tools/quickbench.py 1024 > gpurun_out/plain2_r2i.log 2>&1; echo "plain2 rc=$?"
KERNELS=stream timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage1_ -s 20 -c 4 -f -o gpurun_out/prof_r2i python tools/quickbench.py 1024 > gpurun_out/ncu_full_r2i.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_r2i.log
