"""Latency of the step's one collective: NCCL all_reduce(AVG) of a 1056-float buffer on the current stream (GradientAllReduce's
pattern) and whether torch's symmetric memory rendezvous works on this box.  torchrun --nproc-per-node N tools/dbg_allreduce.py"""
import os, sys, time
import torch, torch.distributed as dist
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
flat = torch.randn(1056, device=dev)
busy = torch.empty(64 << 20, device=dev)
def one(sync):
    w = dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=True); w.wait()
for _ in range(20): one(True)
torch.cuda.synchronize(); dist.barrier()
ts = []
for i in range(200):
    busy.fill_(1.0)   # some work in front, like the backward kernels
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); one(True); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
msg = f"rank {rank}: NCCL all_reduce 4 KB on the current stream: median {ts[100]:.1f} us  p90 {ts[180]:.1f} us  min {ts[0]:.1f} us"
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(2 * 1056, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    msg += f" | symm_mem ok: world {h.world_size} ptrs {len(h.buffer_ptrs)} signal pad {symm.get_signal_pad_size()} B"
except Exception as e:
    msg += f" | symm_mem FAILED: {type(e).__name__}: {str(e)[:200]}"
for r in range(world):
    dist.barrier()
    if r == rank: print(msg, flush=True)
dist.barrier(); dist.destroy_process_group()
