#!/bin/bash
# round 2, call j: stage-2 primitives on the GPU vs the oracle, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage2.py -x -q > gpurun_out/pytest_r2j_stage2.log 2>&1; echo "stage2 rc=$?"; tail -15 gpurun_out/pytest_r2j_stage2.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2j.log
python - <<'PY'
import sys, torch, time
sys.path.insert(0, '.')
from mojo_simdjson_b200 import device, synth
size = 1 << 30
doc = synth.status_array(size)
d_in = torch.from_numpy(doc).cuda(); d_idx = torch.empty(size // 3 + 64, dtype=torch.int32, device='cuda')
ctx = device.Stage1Context(0)
res = ctx.index(d_in, d_idx)
for want_strings in (False, True):
    for _ in range(2):
        out = ctx.stage2_primitives(d_in, d_idx, res.n, want_strings)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ctx.stage2_primitives(d_in, d_idx, res.n, want_strings)
    e1.record(); torch.cuda.synchronize()
    s = out['summary'].cpu().tolist()
    print('stage2 primitives, 1 GiB document, %d structurals, strings written: %s: %.3f ms per pass, first error %d, string bytes %d' % (res.n, want_strings, e0.elapsed_time(e1) / 5, s[0], s[1]))
PY
