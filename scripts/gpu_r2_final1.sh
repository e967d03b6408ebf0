#!/bin/bash
# round 2, final single-GPU measurements of the committed build: all bench configurations, reference arm, size sweep, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_final.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
python __graft_entry__.py smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final_doc1g.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final_doc1g.log | cut -c1-250
for cfg in small cjk dense runs; do
  timeout 600 python bench.py --config $cfg --steps 100 --warmup 10 > gpurun_out/bench_final_$cfg.log 2>&1; echo "bench $cfg rc=$?"; tail -1 gpurun_out/bench_final_$cfg.log | cut -c1-200
done
timeout 300 python bench.py --no-utf8 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_final_noutf8.log 2>&1; echo "noutf8 rc=$?"; tail -1 gpurun_out/bench_final_noutf8.log | cut -c1-160
timeout 300 python bench.py --kernel persistent --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_final_persistent.log 2>&1; echo "persistent rc=$?"; tail -1 gpurun_out/bench_final_persistent.log | cut -c1-160
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_final_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_final_ref.log | cut -c1-200
KERNELS=auto,persistent,split,stream SIZES=1,4,16,32,64,128,256,512,1024 timeout 600 python tools/sizesweep.py > gpurun_out/sizesweep_final.log 2>&1; cat gpurun_out/sizesweep_final.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list_final.log 2>&1; echo "list rc=$?"
KERNELS=stream timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage1_ -s 20 -c 4 -f -o gpurun_out/prof_final python tools/quickbench.py 1024 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/side_outputs.py > gpurun_out/side_outputs_final.log 2>&1; tail -4 gpurun_out/side_outputs_final.log
