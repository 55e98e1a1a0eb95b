"""Host-side cost of one paired step: tiny clouds (the GPU is never the limiter), wall clock per step and per entry point."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from radardistill_b200 import ops, synth
import numpy as np
dev = torch.device("cuda", 0)
lidar = synth.collate([synth.lidar_frame(b, sweeps=1, beams=4, azimuths=64) for b in range(8)])
radar = synth.collate([synth.radar_frame(b, n_points=200) for b in range(8)])
lid, rad, call = bench.build_modules(dev, "B", False)
ld, rd = torch.from_numpy(lidar).to(dev), torch.from_numpy(radar).to(dev)
up = bench.make_upstream(dev, len(lidar), len(radar), list(lid.parameters()) + list(rad.parameters()))
import gc
for _ in range(20):
    bench.gpu_step(call, ld, rd, "B", 8, up)
torch.cuda.synchronize()
gc.collect(); gc.freeze(); gc.disable()
K = 300
ts = []
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(K):
        bench.gpu_step(call, ld, rd, "B", 8, up)
    torch.cuda.synchronize()
    ts.append(1e6 * (time.perf_counter() - t0) / K)
print(f"paired step, tiny clouds ({len(lidar)} + {len(radar)} rows): {min(ts):.0f} us/step wall (host + launch-latency bound)")
acc = {}
def wrap(mod, name):
    fn = getattr(mod, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(mod, name, w)
for n in ("encode_launch", "encode_finish", "encode_backward"):
    wrap(ops, n)
t0 = time.perf_counter()
for _ in range(K):
    bench.gpu_step(call, ld, rd, "B", 8, up)
torch.cuda.synchronize()
print(f"  instrumented: {1e6*(time.perf_counter()-t0)/K:.0f} us/step")
for n, v in acc.items():
    print(f"  {n:16s} {1e6*v/K:7.1f} us/step (2 calls)")
