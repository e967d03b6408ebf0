"""Latency of the verdict exchange alone (all-gather of 16 int32 per rank + all-reduce(MAX) of 16 int32), under torchrun."""
import os, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = torch.zeros(16, dtype=torch.int32, device="cuda"); g = torch.empty(16 * world, dtype=torch.int32, device="cuda")
for _ in range(20):
    dist.all_gather_into_tensor(g, st); dist.all_reduce(st, op=dist.ReduceOp.MAX)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_gather_into_tensor(g, st); dist.all_reduce(st, op=dist.ReduceOp.MAX)
e1.record(); torch.cuda.synchronize()
if rank == 0: print("world", world, "exchange (all_gather + all_reduce) us", e0.elapsed_time(e1) / 200 * 1e3, flush=True)
dist.destroy_process_group()
