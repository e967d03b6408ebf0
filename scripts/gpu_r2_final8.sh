#!/bin/bash
# round 2, final multi-GPU measurements (one 8-GPU box): config 4 at N = 2, 4, 8 through the in-library batch driver; reference arm at N = 8
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/bench_final8_${n}gpu.log 2>&1; echo "bench$n rc=$?"; tail -1 gpurun_out/bench_final8_${n}gpu.log | cut -c1-200
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_final8_1gpu.log 2>&1; echo "bench1 rc=$?"; tail -1 gpurun_out/bench_final8_1gpu.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --impl reference --steps 5 --warmup 2 > gpurun_out/bench_final8_ref.log 2>&1; echo "ref8 rc=$?"
export SJB200_NCCL_LIB=$(python -c "import nvidia.nccl, os; print(os.path.join(list(nvidia.nccl.__path__)[0], 'lib', 'libnccl.so.2'))")
timeout 600 ./tests/cpp/batch_test 8 256 2 > gpurun_out/batch_test_8gpu.log 2>&1; echo "batch_test8 rc=$?"; tail -4 gpurun_out/batch_test_8gpu.log
