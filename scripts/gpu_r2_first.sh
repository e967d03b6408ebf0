#!/bin/bash
# first GPU call of round 2: parity of the fused kernel, then fused vs stream on the 1 GiB document for a few builds
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r2a.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2a.log
for v in "" d3 lag2 nw16 d3lag2; do
  SJB200_LIB_VARIANT=$v KERNELS=fused,stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -2
done
SJB200_LIB_VARIANT= KERNELS=fused,stream,split,persistent timeout 300 python tools/quickbench.py 64 2>&1 | tail -4
SJB200_LIB_VARIANT= KERNELS=fused,stream,split,persistent timeout 300 python tools/quickbench.py 16 2>&1 | tail -4
timeout 600 python bench.py --steps 50 --warmup 10 --e2e-steps 2 > gpurun_out/bench_r2a.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2a.log | cut -c1-1500
