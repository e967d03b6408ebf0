#!/bin/bash
# round 2, call k (2 GPUs): the batch driver from C++ alone on 2 GPUs, bench --gpus 2 through the in-library driver, reference arm at N=2
mkdir -p gpurun_out
export SJB200_NCCL_LIB=$(python -c "import nvidia.nccl, os; print(os.path.join(os.path.dirname(nvidia.nccl.__file__), 'lib', 'libnccl.so.2'))")
timeout 600 ./tests/cpp/batch_test 2 64 2 > gpurun_out/batch_test_2gpu.log 2>&1; echo "batch_test rc=$?"; tail -5 gpurun_out/batch_test_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2k_2gpu.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_r2k_2gpu.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe.py > gpurun_out/pcie_probe_2gpu.log 2>&1; echo "probe rc=$?"; tail -2 gpurun_out/pcie_probe_2gpu.log | cut -c1-900
