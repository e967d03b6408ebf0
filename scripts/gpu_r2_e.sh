#!/bin/bash
# round 2, call e: windowed stream pipeline -- window size x classify occupancy sweep, then parity tests of the stream path
mkdir -p gpurun_out
for cfg in "0 4" "0 2" "32 2" "64 2" "128 2" "64 3" "128 3" "64 4" "256 2" "16 2"; do
  set -- $cfg
  SJB200_WINDOW_MIB=$1 SJB200_CLASSIFY_CTAS=$2 KERNELS=stream timeout 300 python tools/quickbench.py 1024 2>&1 | tail -1 | sed "s/^/win=$1 ctas=$2 /"
done
timeout 1500 python -m pytest tests -m gpu -x -q -k "stream or speculation or dense or heavy or corpus or utf8 or kernel_and_tile" > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2e.log
timeout 600 python bench.py --config runs --steps 50 --warmup 10 > gpurun_out/bench_r2e_runs.log 2>&1; echo "bench runs rc=$?"; tail -1 gpurun_out/bench_r2e_runs.log | cut -c1-200
