"""Kernel-only throughput across document sizes with the library's automatic kernel choice (and forced ones).
usage: python tools/sizesweep.py ; env KERNELS=auto,persistent,split,stream SIZES=1,4,16,64,256,1024 (MiB)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_simdjson_b200 import device, synth
sizes = [int(x) for x in os.environ.get('SIZES', '1,4,16,32,64,128,256,1024').split(',')]
kernels = os.environ.get('KERNELS', 'auto,persistent').split(',')
big = synth.status_array(max(sizes) << 20)
ctx = device.Stage1Context(0)
print('size_mib ' + ' '.join('%12s' % k for k in kernels))
for mib in sizes:
    # a prefix of the big document is not a complete document, but stage 1 does not care: same bytes, same work
    size = mib << 20
    d_in = torch.from_numpy(big[:size]).cuda(); d_out = torch.empty(size // 3 + 16, dtype=torch.int32, device='cuda')
    row = []
    for kernel in kernels:
        ctx.set_kernel(kernel)
        for _ in range(5): ctx.enqueue(d_in, d_out, 0)
        torch.cuda.synchronize()
        reps = max(20, min(400, (8 << 30) // size))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): ctx.enqueue(d_in, d_out, 0)
        e1.record(); torch.cuda.synchronize()
        ctx.finish()
        ms = e0.elapsed_time(e1) / reps
        row.append('%6.0f GB/s' % (size / ms / 1e6) + ' %5.0fus' % (ms * 1e3) if False else '%7.0f/%6.1fus' % (size / ms / 1e6, ms * 1e3))
    print('%8d ' % mib + ' '.join('%12s' % r for r in row), flush=True)
