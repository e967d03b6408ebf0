"""CPU tests (gloo, world_size 2) of the batch driver's host logic: sharding at line starts and the verdict exchange.

No compute happens here -- the per-segment results are canned; the indexing itself only exists on the GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mojo_simdjson_b200 import batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shards_tile_the_batch_and_start_on_lines():
    rng = np.random.default_rng(0)
    lines = [b'{"k":' + str(int(x)).encode() * int(n) + b"}\n" for x, n in zip(rng.integers(0, 99, 500), rng.integers(1, 40, 500))]
    data = np.frombuffer(b"".join(lines), dtype=np.uint8)
    total = data.size
    for world in (1, 2, 3, 4, 8):
        prev_end = 0
        for rank in range(world):
            lo, hi = batch.shard_byte_range(total, world, rank)
            a, b = batch.align_to_lines(data, lo, hi, total)
            assert a == prev_end                      # shards tile the batch: nothing lost, nothing twice
            assert a == 0 or data[a - 1] == 0x0A      # every shard starts on a line start
            assert b == total or data[b - 1] == 0x0A
            prev_end = b
        assert prev_end == total


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # canned per-segment results: rank 1 has a failing segment (UNCLOSED_STRING = 15)
        if rank == 0:
            errors = torch.tensor([0, 0, 0], dtype=torch.int32)
            counts = torch.tensor([100, 200, 300], dtype=torch.int32)
        else:
            errors = torch.tensor([0, 15], dtype=torch.int32)
            counts = torch.tensor([7, 0], dtype=torch.int32)
        worst, all_counts = batch.exchange_verdicts(errors, counts, max_segments=4)
        q.put((rank, int(worst.item()), all_counts.tolist()))
    finally:
        dist.destroy_process_group()


def test_verdict_exchange_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, worst, all_counts in got:
        assert worst == 15                                  # every rank learns the worst verdict
        assert all_counts == [[100, 200, 300, -1], [7, 0, -1, -1]]  # and every rank's per-segment counts


def test_exchange_without_process_group_is_local():
    worst, allc = batch.exchange_verdicts(torch.tensor([0, 13], dtype=torch.int32), torch.tensor([5, 0], dtype=torch.int32), 3)
    assert int(worst.item()) == 13 and allc.tolist() == [[5, 0, -1]]
