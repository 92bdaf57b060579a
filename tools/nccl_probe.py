"""Times the collectives the sharded filter issues (torch.distributed over NCCL, CUDA events,
max over ranks): the tiny allgather behind every scalar exchange, the bin-occupancy allreduce
(int32 sum over 1.45 M words; the round-1 form: 5.8 M words) and a uint8 max of the same bytes.
Run under torchrun; NCCL_DEBUG=INFO shows the transport (NVLS / P2P / SHM)."""
import json, os, sys
import torch, torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def timed(fn, iters=50, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3  # us

out = {"world": world}
small = torch.zeros(world * 8, dtype=torch.uint8, device="cuda")
out["allgather_8B_us"] = timed(lambda: dist.all_gather_into_tensor(small, small[rank * 8:(rank + 1) * 8]))
one = torch.zeros(1, dtype=torch.int64, device="cuda")
out["allreduce_8B_us"] = timed(lambda: dist.all_reduce(one))
for name, n, dt, op in (("allreduce_i32_sum_5.8MB_us", 1_440_000, torch.int32, dist.ReduceOp.SUM),
                        ("allreduce_i32_sum_23MB_us", 5_760_000, torch.int32, dist.ReduceOp.SUM),
                        ("allreduce_u8_max_5.8MB_us", 5_760_000, torch.uint8, dist.ReduceOp.MAX),
                        ("allreduce_i32_sum_360KB_us", 90_000, torch.int32, dist.ReduceOp.SUM)):
    t = torch.zeros(n, dtype=dt, device="cuda")
    out[name] = timed(lambda: dist.all_reduce(t, op=op), iters=30)
# the host-synchronised form of a scalar exchange (shard_gather): allgather + D2H + stream sync
pinned = torch.zeros(world * 8, dtype=torch.uint8).pin_memory()
def gather_sync():
    dist.all_gather_into_tensor(small, small[rank * 8:(rank + 1) * 8])
    pinned.copy_(small, non_blocking=True)
    torch.cuda.current_stream().synchronize()
import time
for _ in range(10): gather_sync()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100): gather_sync()
out["allgather_8B_d2h_sync_wall_us"] = (time.perf_counter() - t0) / 100 * 1e6
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
