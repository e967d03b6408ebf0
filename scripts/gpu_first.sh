#!/bin/bash
# first GPU call: smoke -> parity tests -> bench sweep
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.log
for w in 2 4 8; do
  timeout 300 python bench.py --steps 50 --warmup 10 --warps $w --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_w$w.log 2>&1; echo "bench w$w rc=$?"; tail -1 gpurun_out/bench_w$w.log | cut -c1-400
done
timeout 300 python bench.py --steps 50 --warmup 10 --no-utf8 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_noutf8.log 2>&1; echo "bench noutf8 rc=$?"; tail -1 gpurun_out/bench_noutf8.log | cut -c1-400
