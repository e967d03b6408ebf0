#!/bin/bash
# round 2, final multi-GPU measurement at N GPUs (N = first argument): config 4 through the in-library batch driver
n=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/bench_final8_${n}gpu.log 2>&1; echo "bench$n rc=$?"; tail -1 gpurun_out/bench_final8_${n}gpu.log | cut -c1-200
