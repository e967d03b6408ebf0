"""NDJSON / multi-document batch driver: one process per GPU, `torch.distributed` for the plumbing.

A batch is cut at newlines into independent segments (each is exactly one reference stage-1 call: segment-relative
indexes, own trailer, own verdict).  Ranks own contiguous byte ranges of the batch; segments never straddle ranks.
After a pass the ranks exchange verdicts -- an all-reduce(MAX) of the error flag and an all-gather of the per-segment
counts -- which is the ONLY inter-GPU traffic of this path; bulk data never leaves its GPU.

The sharding / exchange logic is device agnostic (NCCL on GPUs, gloo in the CPU tests); the indexing itself always
runs on the GPU through libsimdjson_b200.so -- there is no CPU implementation in this package.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


def shard_byte_range(total_bytes: int, world: int, rank: int) -> tuple[int, int]:
    """Nominal byte range of `rank` before it is moved to line starts: equal contiguous slices."""
    per = -(-total_bytes // world)
    lo = min(rank * per, total_bytes)
    hi = min(lo + per, total_bytes)
    return lo, hi


def align_to_lines(buf, lo: int, hi: int, total: int) -> tuple[int, int]:
    """Moves both ends of [lo, hi) forward to the next line start (a position right after a '\\n', or 0 / total).

    `buf` is anything indexable by byte position (bytes, numpy uint8 array, 1-D CPU tensor).  Every rank applies the
    same rule to its own ends, so the shards tile the batch exactly and no line is split."""

    def next_line_start(p: int) -> int:
        if p <= 0:
            return 0
        while p < total and int(buf[p - 1]) != 0x0A:
            p += 1
        return p

    return next_line_start(lo), next_line_start(hi) if hi < total else total


@dataclass
class BatchVerdict:
    worst_error: int            # max over every segment of every rank (0 = all fine)
    counts: list[list[int]]     # counts[rank][segment]
    errors: list[int]           # this rank's per-segment error codes
    utf8: list[int]             # this rank's per-segment UTF-8 verdicts


def exchange_verdicts(errors: torch.Tensor, counts: torch.Tensor, max_segments: int, group=None) -> tuple[torch.Tensor, torch.Tensor]:
    """The two collectives of the path.  errors: int32 [nseg] (this rank), counts: int32 [nseg].

    Returns (worst, all_counts): worst = all-reduce(MAX) of the local worst error (int32 [1]); all_counts =
    all-gather of the counts padded to max_segments (int32 [world, max_segments], -1 = no such segment).
    Works on CPU tensors (gloo) and CUDA tensors (NCCL, enqueued behind the kernels: no host round trip)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    worst = errors.max().reshape(1).to(torch.int32) if errors.numel() else torch.zeros(1, dtype=torch.int32, device=errors.device)
    padded = torch.full((max_segments,), -1, dtype=torch.int32, device=counts.device)
    padded[: counts.numel()] = counts
    if world == 1:
        return worst, padded.reshape(1, max_segments)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    gathered = torch.empty(world * max_segments, dtype=torch.int32, device=counts.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    return worst, gathered.reshape(world, max_segments)


class NdjsonBatchDriver:
    """Indexes this rank's shard of an NDJSON batch that is already resident on its GPU."""

    def __init__(self, ctx, seg_bytes: int = 1 << 30, max_segments: int = 64, group=None):
        self.ctx = ctx                      # mojo_simdjson_b200.device.Stage1Context
        self.seg_bytes = min(seg_bytes, 0x7FFFFFFF)
        self.max_segments = max_segments
        self.group = group
        self._status = None               # rows written by the most recent pass
        self._offsets = None
        self._bufs = None
        self._last = None
        self._pass = 0
        # True: the exchange of a pass runs beside the kernels of the next one; False: strictly after its own pass
        self.overlap_exchange = os.environ.get("SJB200_EXCHANGE", "async") != "sync"

    def plan(self, d_shard: torch.Tensor) -> list[int]:
        """Cuts the shard at newlines into segments of < 2 * seg_bytes (device-side search)."""
        self._offsets = self.ctx.split(d_shard, self.seg_bytes, self.max_segments)
        nseg = len(self._offsets) - 1
        if nseg > self.max_segments:
            raise ValueError(f"{nseg} segments > max_segments={self.max_segments}")
        self._status = torch.full((self.max_segments, 2), -1, dtype=torch.int32, device=d_shard.device)
        return self._offsets

    def index_capacity(self) -> int:
        return self._offsets[-1] + 3 * (len(self._offsets) - 1)

    def enqueue(self, d_shard: torch.Tensor, d_out: torch.Tensor, flags: int = 0):
        """One pass: every segment's kernels on the current stream, then the verdict exchange -- an all-gather of the
        per-segment {error, count} rows and an all-reduce(MAX) of the same rows -- issued asynchronously, so that on the
        GPU it runs beside the kernels of the NEXT pass (two sets of buffers alternate; a set is reused only after the
        exchange that read it has finished, a stream-level wait that never blocks the host).

        Returns (errors, all_counts): errors int32 [max_segments] = per segment index the worst error over all ranks
        (-1 = no rank has such a segment), all_counts int32 [world, max_segments] (-1 = no such segment).  Both are valid
        once flush() has been enqueued and the stream synchronised (run() does that)."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if self._bufs is None:
            dev = d_shard.device
            self._bufs = [dict(status=torch.full((self.max_segments, 2), -1, dtype=torch.int32, device=dev),
                               gathered=torch.empty(world * self.max_segments * 2, dtype=torch.int32, device=dev), works=())
                          for _ in range(2)]
        b = self._bufs[self._pass & 1]
        self._pass += 1
        for w in b["works"]:
            w.wait()                       # the exchange of two passes ago must be done with these buffers
        b["works"] = ()
        rc = self.ctx.run_segments_async(d_shard, self._offsets, d_out, b["status"], flags)
        if rc != 0:
            raise RuntimeError(f"segment launch failed with error {rc}")
        self._status = b["status"]
        if world > 1:
            # both collectives run in issue order on the communicator's stream: the gather reads the local rows before the
            # in-place reduction turns them into the maximum over the ranks
            overlap = self.overlap_exchange
            w1 = dist.all_gather_into_tensor(b["gathered"], b["status"].reshape(-1), group=self.group, async_op=overlap)
            w2 = dist.all_reduce(b["status"], op=dist.ReduceOp.MAX, group=self.group, async_op=overlap)
            b["works"] = (w1, w2) if overlap else ()
            allc = b["gathered"].reshape(world, self.max_segments, 2)[:, :, 1]
        else:
            allc = b["status"][:, 1].reshape(1, self.max_segments)
        self._last = b
        return b["status"][:, 0], allc

    def flush(self) -> None:
        """Makes the current stream wait for every verdict exchange still in flight."""
        for b in self._bufs or ():
            for w in b["works"]:
                w.wait()
            b["works"] = ()

    def run(self, d_shard: torch.Tensor, d_out: torch.Tensor, flags: int = 0) -> BatchVerdict:
        errors, all_counts = self.enqueue(d_shard, d_out, flags)
        self.flush()
        torch.cuda.synchronize(d_shard.device)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        nseg = len(self._offsets) - 1
        counts = [[int(c) for c in row if c >= 0] for row in all_counts.cpu().tolist()]
        if world > 1:
            mine = self._last["gathered"].reshape(world, self.max_segments, 2)[rank, :nseg, 0].cpu()
        else:
            mine = self._last["status"][:nseg, 0].cpu()
        return BatchVerdict(int(errors.max().item()), counts, [int(x) for x in mine], [])
