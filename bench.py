#!/usr/bin/env python
"""bench.py -- stage-1 input GB/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (synthetic, deterministic; mojo_simdjson_b200/synth/gen.c):
  N = 1 : BASELINE.json configs[2] -- ONE 1 GiB JSON document (array of minified twitter-like statuses).
  N > 1 : configs[3] weak-scaled -- 1 GiB of NDJSON per GPU (N GiB total, 8 GiB at N = 8), cut at newlines into
          independent segments; after every pass the ranks exchange verdicts with NCCL (all-reduce MAX of the error
          flag, all-gather of the per-segment counts) -- the only inter-GPU traffic of this path.
A step = one stage-1 pass over the rank's input.  `value` times it with the input resident in HBM (CUDA events on
the launching stream, max over ranks); `e2e` times the same pass through the host-buffer C-ABI call
(sjb200_stage1: H2D copy of the document from pinned memory + kernel + D2H copy of the n+3 indexes).
The input (1 GiB) is larger than L2 (126 MB), so no L2 flush is needed between iterations.

--impl reference times the CPU restatement of the reference's stage 1 (oracle/, kind "port": the reference is Mojo
and cannot be built in this image) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stage1_input_throughput"
UNIT = "GB/s"
GIB = 1 << 30
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


def ncu_traffic(kind: str):
    """Committed ncu capture of the kernel organisation `kind` ("stream" | "persistent"), if any: dram bytes per document
    pass (dram__bytes_read.sum + dram__bytes_write.sum over its launches) and the per-kernel shares."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kind) or {}
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, ~4 ms period)."""

    def __init__(self, device_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes-per-gpu", type=int, default=GIB)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--warps", type=int, default=0, help="force tile shape (2/4/8), 0 = auto")
    ap.add_argument("--kernel", default="auto", choices=["auto", "persistent", "split", "stream", "fused"],
                    help="force the kernel organisation (sjb200_ctx_set_kernel); auto = the library's choice")
    ap.add_argument("--no-utf8", action="store_true", help="skip UTF-8 validation (the reference validates nothing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-bytes", type=int, default=GIB)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference) -- the only place besides tests/smoke that runs oracle/
# ------------------------------------------------------------------------------------------------
def cpu_stage1_gbs(sample, threads: int, repeats: int = 1):
    """Times oracle_stage1_ref (block-for-block restatement incl. the 64-iteration prefix_xor) on `sample`,
    split into `threads` independent pieces run concurrently (ctypes releases the GIL)."""
    import numpy as np

    from oracle import oracle

    L = oracle.lib(native=True)
    n_bytes = int(sample.size)
    piece = (n_bytes + threads - 1) // threads
    parts = [sample[i * piece : min(n_bytes, (i + 1) * piece)] for i in range(threads)]
    outs = [np.empty(p.size // 3 + 16, dtype=np.uint32) for p in parts]

    def work(i):
        n = C.c_uint32(0)
        nw = C.c_uint64(0)
        u8 = C.c_int32(0)
        p = parts[i]
        L.oracle_stage1_ref(p.ctypes.data, p.size, outs[i].ctypes.data, outs[i].size, C.byref(n), C.byref(nw), C.byref(u8), 0)

    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        if threads == 1:
            work(0)
        else:
            ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
            [t.start() for t in ts]
            [t.join() for t in ts]
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return n_bytes / best / 1e9, best


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    """Reference arm: rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from mojo_simdjson_b200 import synth

    n_gpus = args.gpus
    cores = os.cpu_count() or 1
    # the reference is a single-threaded, sequential scan: one document can use one core; an NDJSON batch can use
    # one core per independent segment
    threads = 1 if n_gpus == 1 else min(cores, 64)
    per_step = (64 << 20) if n_gpus == 1 else (32 << 20) * threads
    if n_gpus == 1:
        sample = synth.status_array(per_step, synth.SEED_DOC)
    else:
        sample = synth.ndjson(per_step, synth.SEED_NDJSON)
    for _ in range(min(args.warmup, 2)):
        cpu_stage1_gbs(sample, threads)
    t_total = 0.0
    steps = args.steps
    budget_s = 150.0
    done = 0
    for _ in range(steps):
        _, dt = cpu_stage1_gbs(sample, threads)
        t_total += dt
        done += 1
        if t_total > budget_s:
            break
    gbs = per_step * done / t_total / 1e9
    sample_desc = (f"{per_step >> 20} MiB prefix-sized sample of the same generator per step, {done} timed steps, "
                   f"{threads} thread(s); CPU: {cpu_model()}, {cores} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": UNIT, "n_gpus": n_gpus, "steps": done,
        "warmup": min(args.warmup, 2), "ms_per_step": round(1e3 * t_total / done, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(n_gpus, args.bytes_per_gpu),
        "cpu_baseline": {"value": round(gbs, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": round(gbs, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference's stage 1 (oracle/stage1_oracle.c, -O3 -march=native); the Mojo "
                "reference itself cannot be built in this image",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, bytes_per_gpu: int):
    if n_gpus == 1:
        return {"workload": f"single synthetic {bytes_per_gpu / GIB:g} GiB JSON document (array of minified twitter-like "
                            "statuses, seed 0x5EED0002) on 1 B200", "bytes_per_gpu": bytes_per_gpu,
                "l2": "input larger than the 126 MB L2, no flush between iterations"}
    return {"workload": f"{n_gpus} x {bytes_per_gpu / GIB:g} GiB synthetic NDJSON batch (one status object per line, seed "
                        "0x5EED0003+rank) sharded by line ranges, one shard per GPU; NCCL all-reduce(MAX) of error flags + "
                        "all-gather of segment counts per pass", "bytes_per_gpu": bytes_per_gpu,
            "l2": "input larger than the 126 MB L2, no flush between iterations"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from mojo_simdjson_b200 import _native, device, errors, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the stage-1 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    size = args.bytes_per_gpu
    flags = _native.FLAG_NO_UTF8 if args.no_utf8 else 0

    # ---- workload: generated straight into pinned host memory, then resident in HBM -------------------
    h_in = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    if n_gpus == 1:
        synth.status_array(size, synth.SEED_DOC, out=h_in.numpy())
    else:
        synth.ndjson(size, synth.SEED_NDJSON + rank, out=h_in.numpy())
    d_in = h_in.to(dev, non_blocking=True)
    cap = size // 3 + 64
    d_out = torch.empty(cap, dtype=torch.int32, device=dev)
    h_out = torch.empty(cap, dtype=torch.int32, pin_memory=True)
    ctx = device.Stage1Context(local, max_len=(1 << 32) - 1, max_len_host=size)
    stream = torch.cuda.current_stream(dev)
    ctx.use_stream(stream)
    if args.warps:
        ctx.set_warps(args.warps)
    if args.kernel != "auto":
        ctx.set_kernel(args.kernel)
    torch.cuda.synchronize()

    from mojo_simdjson_b200 import batch

    driver = None
    if n_gpus == 1:
        seg_offsets = [0, size]
        nseg = 1
    else:
        driver = batch.NdjsonBatchDriver(ctx, seg_bytes=min(size, 0x7FFFFFFF), max_segments=8)
        seg_offsets = driver.plan(d_in)
        nseg = len(seg_offsets) - 1
    last_exchange = {}

    def step():
        if n_gpus == 1:
            rc = ctx.enqueue(d_in, d_out, flags)
            if rc != errors.SUCCESS:
                raise RuntimeError(f"launch failed: {errors.NAMES.get(rc, rc)}")
        else:
            # every segment's kernels, then the verdict exchange (all-gather of the per-segment rows, all-reduce MAX of
            # the error flags), issued asynchronously behind the kernels: no host round trip, and on the GPU it runs
            # beside the kernels of the next pass
            last_exchange["worst"], last_exchange["counts"] = driver.enqueue(d_in, d_out, flags)

    # ---- correctness gate before any timing -------------------------------------------------------------
    if n_gpus == 1:
        res = ctx.index(d_in, d_out, flags)
        n_total = res.n
        if res.error != 0 or not n_total:
            raise RuntimeError(f"stage 1 failed on the bench document: {res}")
        tr = d_out[n_total : n_total + 3].cpu().numpy().view(np.uint32).tolist()
        assert tr == [size & 0xFFFFFFFF, size & 0xFFFFFFFF, 0], tr
    else:
        step()
        driver.flush()
        torch.cuda.synchronize()
        if int(last_exchange["worst"].max().item()) != 0:
            raise RuntimeError(f"stage 1 failed on an NDJSON segment: {driver._status.cpu().tolist()}")
        allc = last_exchange["counts"].cpu()
        n_total = int(allc[rank][allc[rank] >= 0].sum())
        assert allc.shape[0] == world
    density = n_total / size

    # ---- timed region: K passes, device-resident ----------------------------------------------------------
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)   # NVML initialisation takes tens of ms with 8 processes: before the barrier, not after
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev0.record(stream)
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    if driver is not None:
        driver.flush()   # the last verdict exchanges belong to the timed region
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps   # CPU time to issue one pass (diagnostic)
    ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    launches = ctx.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = float(t.item()) / args.steps
    value = n_gpus * size / (ms_step * 1e-3) / 1e9

    # ---- dominant kernel alone (same launch, CUDA events on its stream) for the roofline -------------------
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kreps = max(10, min(args.steps, 50))
    torch.cuda.synchronize()
    kev0.record(stream)
    for _ in range(kreps):
        if n_gpus == 1:
            ctx.enqueue(d_in, d_out, flags)
        else:
            ctx.run_segments_async(d_in, seg_offsets, d_out, driver._status, flags)
    kev1.record(stream)
    torch.cuda.synchronize()
    k_ms = kev0.elapsed_time(kev1) / kreps / nseg  # per document pass
    if n_gpus == 1:
        # at N = 1 a timed step IS one document pass (same launches, same stream, same events): report the roofline from
        # the timed region itself so that `value` and `roofline.achieved` can never disagree; the loop above is kept as a
        # cross-check
        k_ms_check, k_ms = k_ms, ms_step
    else:
        k_ms_check = k_ms
    alg_bytes = (size + 4 * (n_total + 3 * nseg)) / nseg  # per pass: input read once + every index written once
    peak, peak_kind = measured_peak()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # which kernel organisation the library picks for this document size (capi.cu: SPLIT_MIN_BYTES) unless forced
    seg = size / nseg
    kind = args.kernel if args.kernel != "auto" else ("stream" if seg >= (160 << 20) else "split" if seg >= (48 << 20) else "persistent")
    prof = ncu_traffic(kind)
    if kind == "stream":
        kname = ("stage-1 stream pipeline, 6 launches per document: stage1_stream_classify_kernel -> stage1_utf8_lanes_kernel -> "
                 "stage1_span_reduce_kernel -> stage1_span_carries_kernel -> stage1_flatten_kernel (+ stage1_persistent_kernel as a "
                 "no-op fallback)")
    else:
        kname = {"persistent": "stage1_persistent_kernel", "split": "stage1_classify_kernel + stage1_flatten_kernel",
                 "fused": "stage1_fused_kernel (classify / scan / flatten interleaved in one persistent launch; + stage1_persistent_kernel "
                          "as a no-op fallback)"}[kind]
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": prof.get("dram_bytes_per_pass"), "peak_source": f"of {peak_kind}",
                "kernel": kname, "kernel_ms": round(k_ms, 4), "kernel_ms_second_loop": round(k_ms_check, 4), "algorithmic_bytes_per_launch": int(alg_bytes),
                "structural_density": round(density, 4),
                "note": "achieved = algorithmic bytes of one document pass / device time of ALL its launches (CUDA events)",
                "ncu_kernel_shares": prof.get("kernel_shares")}

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------------
    L = _native.lib()
    n_out = C.c_uint32(0)
    u8 = C.c_int32(0)

    h_status = torch.zeros((nseg, 2), dtype=torch.int32)

    def e2e_step():
        worst = 0
        for s in range(nseg):
            a, b = seg_offsets[s], seg_offsets[s + 1]
            rc = L.sjb200_stage1(ctx._ctx, h_in.data_ptr() + a, b - a, h_out.data_ptr(), cap, C.byref(n_out), C.byref(u8), flags)
            h_status[s, 0] = rc
            h_status[s, 1] = n_out.value
            worst = max(worst, rc)
        if world > 1:
            d_st = h_status.to(dev, non_blocking=True)
            batch.exchange_verdicts(d_st[:, 0], d_st[:, 1], driver.max_segments)
        return worst

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    eev0, eev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eev0.record(stream)
    for _ in range(args.e2e_steps):
        rc = e2e_step()
    eev1.record(stream)
    torch.cuda.synchronize()
    e_ms = torch.tensor([eev0.elapsed_time(eev1) / args.e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * size / (float(e_ms.item()) * 1e-3) / 1e9
    got = h_out[: n_out.value + 3].numpy().view(np.uint32) if nseg == 1 else None
    e2e = {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(size),
           "d2h_bytes_per_step": int(4 * (n_total + 3 * nseg)), "steps": args.e2e_steps,
           "ms_per_step": round(float(e_ms.item()), 3), "api": "sjb200_stage1 (pinned host in, pinned host out)"}
    if rc != 0:
        raise RuntimeError(f"e2e pass failed: {rc}")
    if got is not None:
        assert int(got[-3]) == size % (1 << 32) and int(got[-1]) == 0 and n_out.value == n_total

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        sample_bytes = min(size, args.cpu_sample_bytes)
        sample = h_in.numpy()[:sample_bytes]
        gbs, secs = cpu_stage1_gbs(sample, 1)
        # bit-exact spot check of the GPU stream against the oracle on that sample's prefix
        from oracle import oracle

        chk = min(sample_bytes, 64 << 20)
        want = oracle.stage1(h_in.numpy()[:chk], impl="fast", cap=chk // 3 + 16, native=True)
        k = want.n_written - 1  # the prefix is cut mid-document: compare the indexes that lie inside it
        mine = d_out[:k].cpu().numpy().view(np.uint32)
        parity = bool(np.array_equal(mine, want.indexes[:k]))
        if not parity:
            raise RuntimeError("GPU index stream differs from the oracle on the bench document")
        cpu = {"value": round(gbs, 4), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {sample_bytes >> 20} MiB of the same document, one pass, {secs:.1f} s; oracle_stage1_ref "
                         f"(restatement of the reference incl. its 64-iteration prefix_xor), gcc -O3 -march=native; "
                         f"CPU: {cpu_model()}, {os.cpu_count()} cores",
               "parity_checked_indexes": int(k)}

    # per-rank picture (N > 1): every rank's own pass time without the exchange, own timed-region time and own clocks --
    # the job-level number is the maximum over ranks, so one slower GPU shows up here
    per_rank = None
    if world > 1:
        mine = {"kernel_ms": round(k_ms * nseg, 4), "timed_ms_per_step": round(ms_total / args.steps, 4),
                "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(n_gpus, size),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "utf8_validation": not args.no_utf8, "segments_per_gpu": nseg, "host_enqueue_ms_per_step": round(host_enqueue_ms, 4), "per_rank": per_rank,
            "frac_of_aggregate_hbm": round(value * (alg_bytes * nseg / size) / (peak * n_gpus), 4),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "0"))
    if args.gpus > 1 and world == 0:
        # plain `python bench.py --gpus N`: start one rank per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
