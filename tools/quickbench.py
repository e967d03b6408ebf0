import os, sys, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_simdjson_b200 import device, synth
size = 1 << 30
doc = synth.status_array(size)
d_in = torch.from_numpy(doc).cuda(); d_out = torch.empty(size // 3, dtype=torch.int32, device='cuda')
ctx = device.Stage1Context(0)
for nw in [int(x) for x in os.environ.get('NWS', '8,16,24').split(',')]:
    ctx.set_warps(nw)
    for _ in range(5): ctx.enqueue(d_in, d_out, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): ctx.enqueue(d_in, d_out, 0)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(os.environ.get('SJB200_LIB_VARIANT', 'default'), 'NW', nw, 'ms %.4f' % ms, 'GB/s %.0f' % (size / ms / 1e6))
