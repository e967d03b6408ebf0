#!/bin/bash
# round 2, call l (8 GPUs): config 4 at N = 8 and N = 4 through the in-library batch driver, the reference arm at N = 8, and the
# host <-> device copy ceiling with 1 / 2 / 4 / 8 ranks active
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/bench_r2l_8gpu.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_r2l_8gpu.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 30 --warmup 10 > gpurun_out/bench_r2l_4gpu.log 2>&1; echo "bench4 rc=$?"; tail -1 gpurun_out/bench_r2l_4gpu.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/pcie_probe.py > gpurun_out/pcie_probe_8gpu.log 2>&1; echo "probe rc=$?"; grep -o '"active_ranks": [0-9]*\|"aggregate_gbs": [0-9.]*' gpurun_out/pcie_probe_8gpu.log | paste - - - - -
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 tools/pcie_probe.py --chunk-mib 128 > gpurun_out/pcie_probe_8gpu_c128.log 2>&1; echo "probe128 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 8 --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r2l_8gpu_ref.log 2>&1; echo "ref8 rc=$?"; tail -1 gpurun_out/bench_r2l_8gpu_ref.log | cut -c1-300
