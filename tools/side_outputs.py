"""Timing of the stage-2 side outputs (structural bytes, document starts) on the 1 GiB document.  usage: python tools/side_outputs.py [mib]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_simdjson_b200 import device, synth
size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
doc = synth.status_array(size)
d_in = torch.from_numpy(doc).cuda(); d_idx = torch.empty(size // 3, dtype=torch.int32, device='cuda')
ctx = device.Stage1Context(0)
res = ctx.index(d_in, d_idx)
assert res.error == 0
d_b = torch.empty(res.n, dtype=torch.uint8, device='cuda')
for _ in range(3): ctx.structural_bytes(d_in, d_idx, res.n, d_b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ctx.structural_bytes(d_in, d_idx, res.n, d_b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
moved = size + 5 * res.n
print('structural bytes: n', res.n, 'ms %.4f' % ms, 'input GB/s %.0f' % (size / ms / 1e6), 'algorithmic GB/s (input + 4n + n) %.0f' % (moved / ms / 1e6))

d_s = torch.empty(res.n, dtype=torch.uint8, device='cuda')
for _ in range(3): ctx.document_starts(d_b, res.n, d_s)
torch.cuda.synchronize()
e0.record()
for _ in range(20): ctx.document_starts(d_b, res.n, d_s)
e1.record(); torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / 20
print('document starts: ms %.4f' % ms2, 'GB/s over n bytes in + n bytes out (+ n read again) %.0f' % (3 * res.n / ms2 / 1e6), 'starts', int(d_s.sum().item()))
