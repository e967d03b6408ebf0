#!/bin/bash
# round 2, call c: full GPU test suite, default bench, fused-kernel ablations + one ncu capture of it
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2c.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2c.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2c.log | cut -c1-600
for v in "" abl1 abl2 abl3 abl5; do
  SJB200_LIB_VARIANT=$v KERNELS=fused timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
KERNELS=fused timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage1_fused -s 3 -c 1 -o gpurun_out/prof_fused1 -f python tools/quickbench.py 1024 > gpurun_out/ncu_full_fused1.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full_fused1.log
