"""Two documents indexed concurrently on two streams vs back to back on one: does the flatten kernel of one document
overlap the classify kernel of the other?  usage: python tools/concurrent_docs.py [size_mib]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_simdjson_b200 import device, synth
size = (int(sys.argv[1]) if len(sys.argv) > 1 else 512) << 20
doc = synth.status_array(size)
d_in = torch.from_numpy(doc).cuda()
outs = [torch.empty(size // 3, dtype=torch.int32, device='cuda') for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
ctxs = []
for s in streams:
    c = device.Stage1Context(0)
    c.use_stream(s)
    c.set_kernel('stream')
    ctxs.append(c)
def run(nstreams, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams[:nstreams]: s.wait_event(e0)
    for r in range(reps):
        k = r % nstreams
        ctxs[k].enqueue(d_in, outs[k], 0)
    done = []
    for s in streams[:nstreams]:
        ev = torch.cuda.Event(); ev.record(s); done.append(ev)
    for ev in done: torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (1, 2):
    run(n, 6)
    ms = run(n, 40)
    print('occ', os.environ.get('SJB200_STREAM_OCC', 'max'), 'streams', n, 'ms/doc %.4f' % ms, 'GB/s %.0f' % (size / ms / 1e6), [c.finish().error for c in ctxs], flush=True)
