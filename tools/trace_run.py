"""Debug tool: run one 1 GiB launch of a -DSJ_TRACE=1 build and summarise the per-tile timeline (gpurun box only)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mojo_simdjson_b200 import _native, device, synth

size = 1 << 30
doc = synth.status_array(size)
d_in = torch.from_numpy(doc).cuda()
d_out = torch.empty(size // 3, dtype=torch.int32, device="cuda")
ctx = device.Stage1Context(0)
NW = int(os.environ.get('NW', '8'))
ctx.set_warps(NW)
for _ in range(3):
    res = ctx.index(d_in, d_out)
assert res.error == 0
ntiles = size // (2048 * NW)
L = C.CDLL(_native.LIB_PATH)
L.sjb200_debug_trace.restype = C.c_int32
L.sjb200_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
tr = np.zeros((ntiles, 16), dtype=np.uint64)
rc = L.sjb200_debug_trace(ctx._ctx, tr.ctypes.data, tr.size)
assert rc == 0, rc
t = tr.astype(np.int64)
sel = slice(ntiles // 20, ntiles - ntiles // 20)
def st(name, x):
    x = x[sel]
    print(f"{name:34s} mean {x.mean():9.1f}  p10 {np.percentile(x,10):8.0f}  p50 {np.percentile(x,50):8.0f}  p90 {np.percentile(x,90):8.0f}  p99 {np.percentile(x,99):8.0f} ns")
t0 = t[:, 0].min()
print("kernel span us", (t[:, 13].max() - t0) / 1e3)
st("ticket -> phase1 start (w0)", t[:, 1] - t[:, 0])
per_cta = np.sort(t[:, 1])
st("phase1 start w0 -> w-last", t[:, 7] - t[:, 1])
st("input wait (w0)", t[:, 5])
st("input wait (w-last)", t[:, 6])
st("phase1 start w0 -> agg published", t[:, 2] - t[:, 1])
st("ticket -> agg published", t[:, 2] - t[:, 0])
st("agg published -> lookback start", t[:, 3] - t[:, 2])
st("lookback duration", t[:, 4] - t[:, 3])
st("ring-full wait (w0)", t[:, 14])
st("ring-full wait (w-last)", t[:, 15])
st("carry wait (w0)", t[:, 8])
st("carry wait (w7)", t[:, 9])
st("lookback done -> flush start (w0)", t[:, 10] - t[:, 4])
st("flush duration (w0)", t[:, 12] - t[:, 10])
st("flush duration (w7)", t[:, 13] - t[:, 11])
st("ticket -> flush done (w7)", t[:, 13] - t[:, 0])
# how late is the predecessor's aggregate relative to ours
st("agg(t-1) - agg(t)", np.concatenate([[0], t[:-1, 2] - t[1:, 2]]))
mx = np.maximum.accumulate(t[:, 2])
st("max agg(<t) - agg(t)", np.concatenate([[0], mx[:-1] - t[1:, 2]]))
np.save(os.path.join("gpurun_out", "trace.npy"), tr[:20000])
