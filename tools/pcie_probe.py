"""Host <-> device copy bandwidth with 1 .. N ranks active at once -- the platform ceiling of the end-to-end (host buffer) path.

    python -m torch.distributed.run --nproc-per-node N tools/pcie_probe.py [--mib 1024] [--chunk-mib 32]

Every rank owns one GPU, a pinned input buffer and a pinned output buffer.  For each active-rank count A in 1, 2, 4, .. N
the first A ranks run, simultaneously: H2D alone, D2H alone, both directions on two streams (what sjb200_stage1 does per
document: 1 GiB in, 0.69 GB out), and both directions in `chunk` pieces with an event between them (its pipeline).
Rank 0 prints one JSON line per A with per-rank and aggregate GB/s, plus the NUMA / affinity facts of the box."""
import argparse
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def timed(fn, dev, reps):
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(dev)
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--out-frac", type=float, default=0.645, help="device->host bytes per input byte (4 * 0.161 structurals per byte)")
    ap.add_argument("--chunk-mib", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")
    n_in = args.mib << 20
    n_out = int(n_in * args.out_frac) & ~15
    h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True).fill_(1)
    h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True).fill_(2)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n_out, dtype=torch.uint8, device=dev).fill_(3)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    chunk = args.chunk_mib << 20

    def h2d():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    def pipelined():   # chunked in, chunked out behind an event per chunk (the shape of the streaming host path)
        nchunks = (n_in + chunk - 1) // chunk
        oc = (n_out + nchunks - 1) // nchunks
        for k in range(nchunks):
            a, b = k * chunk, min(n_in, (k + 1) * chunk)
            with torch.cuda.stream(s_in):
                d_in[a:b].copy_(h_in[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev)
                oa, ob = k * oc, min(n_out, (k + 1) * oc)
                if ob > oa:
                    h_out[oa:ob].copy_(d_out[oa:ob], non_blocking=True)

    facts = {}
    if rank == 0:
        for name, cmd in (("numa", "lscpu | grep -i -E 'numa|socket|model name|^cpu\\(s\\)'"), ("topo", "nvidia-smi topo -m | head -14"),
                          ("mem", "grep -E 'MemTotal|MemFree' /proc/meminfo")):
            try:
                facts[name] = subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip().splitlines()
            except Exception as e:
                facts[name] = [repr(e)]
        facts["affinity_rank0"] = sorted(os.sched_getaffinity(0))
    actives = [a for a in (1, 2, 4, 8, 16) if a <= world]
    for fn in (h2d, d2h, both):   # warm up (first pinned copies are slower)
        fn()
    torch.cuda.synchronize(dev)
    for A in actives:
        res = {}
        for name, fn, nbytes in (("h2d", h2d, n_in), ("d2h", d2h, n_out), ("both", both, n_in), ("pipelined", pipelined, n_in)):
            if world > 1:
                dist.barrier()
            dt = timed(fn, dev, args.reps) if rank < A else None
            res[name] = None if dt is None else nbytes / dt / 1e9
            if world > 1:
                dist.barrier()
        allres = [None] * world
        if world > 1:
            dist.all_gather_object(allres, res)
        else:
            allres = [res]
        if rank == 0:
            line = {"active_ranks": A, "world": world, "input_mib": args.mib, "output_bytes_per_input_byte": args.out_frac}
            for name in ("h2d", "d2h", "both", "pipelined"):
                vals = [r[name] for r in allres[:A]]
                line[name] = {"per_rank_gbs": [round(v, 1) for v in vals], "aggregate_gbs": round(sum(vals), 1),
                              "unit": "GB/s of " + ("output" if name == "d2h" else "input") + " bytes; 'both' / 'pipelined' move the output at the same time"}
            if A == actives[0]:
                line["box"] = facts
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
