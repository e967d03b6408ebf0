#!/bin/bash
# verification of the committed build: parity suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/verify_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/verify_pytest.log
python __graft_entry__.py smoke > gpurun_out/verify_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/verify_smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/verify_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/verify_bench.log | cut -c1-220
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/verify_bench_20.log 2>&1; echo "bench20 rc=$?"; tail -1 gpurun_out/verify_bench_20.log | cut -c1-220
