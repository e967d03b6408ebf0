#!/bin/bash
# round 2, final multi-GPU measurement (one 8-GPU box, reduced: 8 GPUs and 1 GPU of the same box, C++ batch test)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_final8_8gpu.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_final8_8gpu.log | cut -c1-200
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_final8_1gpu.log 2>&1; echo "bench1 rc=$?"; tail -1 gpurun_out/bench_final8_1gpu.log | cut -c1-200
export SJB200_NCCL_LIB=$(python -c "import nvidia.nccl, os; print(os.path.join(list(nvidia.nccl.__path__)[0], 'lib', 'libnccl.so.2'))")
timeout 300 ./tests/cpp/batch_test 8 256 2 > gpurun_out/batch_test_8gpu.log 2>&1; echo "batch_test8 rc=$?"; tail -3 gpurun_out/batch_test_8gpu.log
