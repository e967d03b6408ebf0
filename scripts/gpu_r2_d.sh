#!/bin/bash
# round 2, call d: full GPU test suite (in-library batch driver, C++ test), the bench configurations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2d.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2d.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r2d.log | cut -c1-300
for cfg in small cjk dense runs; do
  timeout 600 python bench.py --config $cfg --steps 50 --warmup 10 > gpurun_out/bench_r2d_$cfg.log 2>&1; echo "bench $cfg rc=$?"; tail -1 gpurun_out/bench_r2d_$cfg.log | cut -c1-400
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r2d_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_r2d_ref.log | cut -c1-300
