"""NDJSON / multi-document batch driver: one process per GPU, `torch.distributed` for the plumbing.

A batch is cut at newlines into independent segments (each is exactly one reference stage-1 call: segment-relative
indexes, own trailer, own verdict).  Ranks own contiguous byte ranges of the batch; segments never straddle ranks.
After a pass the ranks exchange verdicts -- ONE all-gather of the per-segment {error, n} rows; the worst error is the
maximum over the gathered rows -- which is the ONLY inter-GPU traffic of this path; bulk data
never leaves its GPU.  Sharding, segment planning, the NCCL communicator and the exchange live inside
libsimdjson_b200.so (csrc/batch_driver.cuh); this module is the Python binding of that driver plus the host-side
sharding helpers, which are device agnostic and covered by gloo tests.  The indexing itself always runs on the GPU --
there is no CPU implementation in this package.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _native, errors


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
    utf8: list[int]             # (per-segment UTF-8 verdicts travel in the error code when SJB200_FLAG_VALIDATE_UTF8 is set)


def exchange_verdicts(errors: torch.Tensor, counts: torch.Tensor, max_segments: int, group=None) -> tuple[torch.Tensor, torch.Tensor]:
    """The verdict exchange of the path as ONE collective, written with torch.distributed: an all-gather of every rank's
    {error, n} rows padded to max_segments.  errors: int32 [nseg] (this rank), counts: int32 [nseg].

    Returns (worst, all_counts): worst = the maximum error over the gathered rows (int32 [1]; no separate all-reduce),
    all_counts = int32 [world, max_segments] (-1 = no such segment).  This is the host-side model of what
    libsimdjson_b200.so does with NCCL inside sjb200_batch_run_resident_async (csrc/batch_driver.cuh); it serves the
    end-to-end (host buffer) path of bench.py and the gloo tests."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rows = torch.full((max_segments, 2), -1, dtype=torch.int32, device=counts.device)
    rows[: errors.numel(), 0] = errors
    rows[: counts.numel(), 1] = counts
    if world == 1:
        gathered = rows.reshape(1, max_segments, 2)
    else:
        flat = torch.empty(world * max_segments * 2, dtype=torch.int32, device=counts.device)
        dist.all_gather_into_tensor(flat, rows.reshape(-1), group=group)
        gathered = flat.reshape(world, max_segments, 2)
    worst = gathered[:, :, 0].max().clamp(min=0).reshape(1).to(torch.int32)
    return worst, gathered[:, :, 1]


class NdjsonBatchDriver:
    """This rank's shard of an NDJSON batch, resident on its GPU, through the in-library batch driver
    (include/simdjson_b200.h: sjb200_batch_create_rank / plan_resident / run_resident_async / finish).

    One process per GPU: the NCCL communicator lives inside libsimdjson_b200.so; torch.distributed is used only to carry
    rank 0's 128-byte NCCL unique id to the other ranks when the driver is created."""

    def __init__(self, device_index: int = 0, seg_bytes: int = 1 << 30, max_segments: int = 64, group=None, stream=None):
        self._lib = _native.lib()
        self.seg_bytes = min(seg_bytes, 0x7FFFFFFF)
        self.max_segments = max_segments
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", device_index)
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                buf = (C.c_uint8 * 128)()
                rc = self._lib.sjb200_batch_unique_id(buf)
                if rc != errors.SUCCESS:
                    raise RuntimeError(f"sjb200_batch_unique_id failed: {errors.NAMES.get(rc, rc)}")
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            carrier = uid.to(self.device) if dist.get_backend(group) == "nccl" else uid
            dist.broadcast(carrier, src=0, group=group)
            uid = carrier.cpu()
        self._uid = (C.c_uint8 * 128)(*uid.tolist())
        self._b = C.c_void_p()
        rc = self._lib.sjb200_batch_create_rank(device_index, self.rank, self.world, self._uid, 0, self.seg_bytes, max_segments, 0,
                                                C.byref(self._b))
        if rc != errors.SUCCESS:
            raise RuntimeError(f"sjb200_batch_create_rank failed: {errors.NAMES.get(rc, rc)}")
        ctx = C.c_void_p()
        self._lib.sjb200_batch_ctx(self._b, 0, C.byref(ctx))
        from . import device as _device

        self.ctx = _device.Stage1Context.borrow(ctx, self.device)   # the driver's own context (kernel choice, stream, launch count)
        self.ctx.use_stream(stream if stream is not None else torch.cuda.current_stream(self.device))
        self._offsets = None
        self._idx_offsets = None
        self._shard = None

    def close(self):
        if self._b:
            self._lib.sjb200_batch_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan(self, d_shard: torch.Tensor) -> list[int]:
        """Cuts the resident shard at newlines into segments of < 2 * seg_bytes (device-side search)."""
        assert d_shard.is_cuda and d_shard.dtype == torch.uint8 and d_shard.is_contiguous()
        offs = (C.c_uint64 * (self.max_segments + 1))()
        ioffs = (C.c_uint64 * (self.max_segments + 1))()
        nseg = C.c_uint32(0)
        rc = self._lib.sjb200_batch_plan_resident(self._b, 0, d_shard.data_ptr(), d_shard.numel(), offs, ioffs, C.byref(nseg))
        if rc != errors.SUCCESS:
            raise RuntimeError(f"sjb200_batch_plan_resident failed: {errors.NAMES.get(rc, rc)}")
        self._shard = d_shard   # keeps the tensor alive while the library holds its pointer
        self._offsets = [int(offs[i]) for i in range(nseg.value + 1)]
        self._idx_offsets = [int(ioffs[i]) for i in range(nseg.value + 1)]
        return self._offsets

    def index_capacity(self) -> int:
        return self._idx_offsets[-1]

    def index_offsets(self) -> list[int]:
        """Entry of the output array where each segment's indexes start (segment s owns len(s) + 3 entries)."""
        return self._idx_offsets

    def enqueue(self, d_out: torch.Tensor, flags: int = 0) -> None:
        """One pass: every segment's kernels on the context's stream, then the all-gather of the {error, n} rows on the
        library's exchange stream (beside the kernels of the next pass).  Never blocks the host."""
        assert d_out.is_cuda and d_out.element_size() == 4 and d_out.is_contiguous()
        ptrs = (C.c_void_p * 1)(d_out.data_ptr())
        caps = (C.c_uint64 * 1)(d_out.numel())
        rc = self._lib.sjb200_batch_run_resident_async(self._b, ptrs, caps, flags)
        if rc != errors.SUCCESS:
            raise RuntimeError(f"sjb200_batch_run_resident_async failed: {errors.NAMES.get(rc, rc)}")

    def finish(self) -> tuple[int, torch.Tensor]:
        """Waits for everything in flight.  Returns (worst error, rows int32 [world, max_segments, 2] of the last pass)."""
        rows = torch.empty((self.world, self.max_segments, 2), dtype=torch.int32)
        worst = C.c_int32(0)
        rc = self._lib.sjb200_batch_finish(self._b, C.cast(rows.data_ptr(), C.POINTER(C.c_int32)), C.byref(worst))
        if rc != errors.SUCCESS:
            raise RuntimeError(f"sjb200_batch_finish failed: {errors.NAMES.get(rc, rc)}")
        return int(worst.value), rows

    def run(self, d_out: torch.Tensor, flags: int = 0) -> BatchVerdict:
        self.enqueue(d_out, flags)
        worst, rows = self.finish()
        nseg = len(self._offsets) - 1
        counts = [[int(n) for e, n in row.tolist() if e >= 0] for row in rows]
        mine = [int(e) for e in rows[self.rank, :nseg, 0].tolist()]
        return BatchVerdict(worst, counts, mine, [])
