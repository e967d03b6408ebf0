#!/bin/bash
# round 2, call m: the stage-2 walk on the GPU (verdict + tape) against the oracle
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage2.py -x -q > gpurun_out/pytest_r2m_stage2.log 2>&1; echo "stage2 rc=$?"; tail -25 gpurun_out/pytest_r2m_stage2.log
