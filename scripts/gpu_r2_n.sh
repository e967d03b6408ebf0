#!/bin/bash
# round 2, call n: facade stage1 + stage2, stage-2 timings at 1 GiB
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage2.py -x -q > gpurun_out/pytest_r2n_stage2.log 2>&1; echo "stage2 rc=$?"; tail -25 gpurun_out/pytest_r2n_stage2.log
python - <<'PY' 2>&1 | tee gpurun_out/stage2_timing_r2n.log
import sys, torch
sys.path.insert(0, '.')
from mojo_simdjson_b200 import device, synth
for mib in (64, 1024):
    size = mib << 20
    doc = synth.status_array(size)
    d_in = torch.from_numpy(doc).cuda(); d_idx = torch.empty(size // 3 + 64, dtype=torch.int32, device='cuda')
    ctx = device.Stage1Context(0)
    res = ctx.index(d_in, d_idx)
    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): out = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out
    ms_p, prims = t(lambda: ctx.stage2_primitives(d_in, d_idx, res.n, True))
    ms_t, (tape, summary) = t(lambda: ctx.stage2_tape(d_in, d_idx, res.n, prims))
    s = summary.cpu().tolist()
    print('%d MiB document, %d structurals: primitives + strings %.3f ms, walk + tape %.3f ms (%.1f GB/s of input for stage 2), verdict %d, tape words %d, inexact doubles %d'
          % (mib, res.n, ms_p, ms_t, size / (ms_p + ms_t) / 1e6, s[0], s[2], s[3]))
    del prims, tape, d_in, d_idx
    ctx.close(); torch.cuda.empty_cache()
PY
