#!/bin/bash
# round 2, call h: merged scan kernel; streaming vs write-back mask stores; size sweep; full GPU tests; the 'runs' document
mkdir -p gpurun_out
for v in "" wb; do
  SJB200_LIB_VARIANT=$v KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1
done
KERNELS=auto,persistent,split,stream SIZES=1,4,16,32,64,128,256,512 timeout 600 python tools/sizesweep.py > gpurun_out/sizesweep_r2h.log 2>&1; cat gpurun_out/sizesweep_r2h.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2h.log
timeout 600 python bench.py --config runs --steps 50 --warmup 10 > gpurun_out/bench_r2h_runs.log 2>&1; echo "bench runs rc=$?"; tail -1 gpurun_out/bench_r2h_runs.log | cut -c1-200
