"""Where does a multi-rank step lose time?  torchrun --nproc-per-node N tools/dbg_scale.py : per-rank host timestamps of
the phases of bench.gpu_step (mode B, lean reducer) + device step times; prints medians / p90 / max per rank."""
import os, sys, time, gc, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
import bench
if world > 1 and os.environ.get("PIN", "1") == "1":
    bench.pin_cpus(local, world)
import numpy as np, torch, torch.distributed as dist
from radardistill_b200 import ops, sharding
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lidar, radar = bench.make_clouds(rank, 8)
lid, rad, call = bench.build_modules(dev, "B", False)
ld, rd = torch.from_numpy(lidar).to(dev), torch.from_numpy(radar).to(dev)
up = bench.make_upstream(dev, len(lidar), len(radar), list(lid.parameters()) + list(rad.parameters()))
red = sharding.GradientAllReduce(up["params"]) if world > 1 else None
T = {}
def wrap(mod, name, key):
    fn = getattr(mod, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); T.setdefault(key, []).append(time.perf_counter() - t); return r
    setattr(mod, name, w)
wrap(ops, "encode_launch", "launch"); wrap(ops, "encode_finish", "finish"); wrap(ops, "encode_backward", "bwd_launch")
def step():
    t0 = time.perf_counter()
    for p in up["params"]: p.grad = None
    bd = call({"points": ld, "radar_points": rd, "batch_size": 8})
    t1 = time.perf_counter()
    outs = [bd["radar_pillar_features"], bd["pillar_features"]]
    torch.autograd.backward(outs, [up["radar"][:outs[0].shape[0]], up["lidar"][:outs[1].shape[0]]])
    t2 = time.perf_counter()
    if red is not None: red.reduce()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for _ in range(10): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
gc.collect(); gc.freeze(); gc.disable()
T.clear()
K = 200
ph, evs = [], []
t_all0 = time.perf_counter()
for _ in range(K):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); ph.append(step()); e.record(); evs.append((s, e))
torch.cuda.synchronize()
wall = (time.perf_counter() - t_all0) / K
gpu = np.array([s.elapsed_time(e) * 1e3 for s, e in evs])
ph = np.array(ph) * 1e6
q = lambda a: "%6.0f %6.0f %6.0f" % (np.median(a), np.percentile(a, 90), np.max(a))
out = f"rank {rank}: wall/step {wall*1e6:.0f} us | device step med/p90/max {q(gpu)} | fwd {q(ph[:,0])} | bwd {q(ph[:,1])} | reduce {q(ph[:,2])}"
for k, v in T.items():
    v = np.array(v) * 1e6
    out += f" | {k} {q(v)}"
for r in range(world):
    if world > 1: dist.barrier()
    if r == rank: print(out, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
