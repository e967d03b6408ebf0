#!/bin/bash
# round 2, call t (re-entry): baseline of the restored build + full ncu capture with per-instruction counts
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/t_bench.log 2>&1
tail -1 gpurun_out/t_bench.log | cut -c1-400
CMD="python bench.py --steps 3 --warmup 1 --no-cpu-baseline --e2e-steps 1"
ncu --set full --clock-control none --import-source on -k regex:stage1_ -s 3 -c 3 -f -o gpurun_out/prof_t $CMD > gpurun_out/t_ncu_full.log 2>&1
echo "profile rc=$?"; tail -3 gpurun_out/t_ncu_full.log
ls -la gpurun_out
