"""Latency of BASELINE.json config 2 (631,515-byte twitter-like document), device-resident and host-to-host."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_simdjson_b200 import _native, device, synth

doc = synth.twitter_like()
d_in = torch.from_numpy(doc).cuda()
d_out = torch.empty(doc.size + 16, dtype=torch.int32, device="cuda")
ctx = device.Stage1Context(0, max_len=1 << 24, max_len_host=1 << 24)
for warps in [int(x) for x in os.environ.get("NWS", "0,2,4,8").split(",")]:
    ctx.set_warps(warps)
    for _ in range(20):
        r = ctx.index(d_in, d_out)
    assert r.error == 0
    # kernel-only: back-to-back launches between two events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 500
    e0.record()
    for _ in range(n):
        ctx.enqueue(d_in, d_out, 0)
    e1.record(); torch.cuda.synchronize()
    k_us = e0.elapsed_time(e1) * 1e3 / n
    # synchronous device-resident call (launch + kernel + verdict visible on the host)
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.index(d_in, d_out)
    s_us = (time.perf_counter() - t0) * 1e6 / n
    print(f"warps={warps}: back-to-back {k_us:.2f} us/launch ({doc.size / k_us / 1e3:.1f} GB/s), synchronous call {s_us:.2f} us ({doc.size / s_us / 1e3:.1f} GB/s)")
# host to host (pinned)
L = _native.lib()
h_in = torch.from_numpy(doc).pin_memory(); h_out = torch.empty(doc.size + 16, dtype=torch.int32).pin_memory()
n_out = C.c_uint32(0); u8 = C.c_int32(0)
ctx.set_warps(0)
for _ in range(20): L.sjb200_stage1(ctx._ctx, h_in.data_ptr(), doc.size, h_out.data_ptr(), h_out.numel(), C.byref(n_out), C.byref(u8), 0)
t0 = time.perf_counter()
for _ in range(300): rc = L.sjb200_stage1(ctx._ctx, h_in.data_ptr(), doc.size, h_out.data_ptr(), h_out.numel(), C.byref(n_out), C.byref(u8), 0)
h_us = (time.perf_counter() - t0) * 1e6 / 300
print(f"host-to-host sjb200_stage1: {h_us:.1f} us ({doc.size / h_us / 1e3:.2f} GB/s), n={n_out.value}, rc={rc}")
