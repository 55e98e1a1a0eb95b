"""Where does the host time of a bench step go?  cProfile over many steps (mode B, no L2 flush), plus wall-clock of
encode_backward / encode_launch / encode_finish measured without the profiler."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from radardistill_b200 import ops
lidar, radar = bench.make_clouds(0, 8)
dev = torch.device("cuda", 0)
lid, rad, call = bench.build_modules(dev, "B", False)
ld, rd = torch.from_numpy(lidar).to(dev), torch.from_numpy(radar).to(dev)
up = bench.make_upstream(dev, len(lidar), len(radar), list(lid.parameters()) + list(rad.parameters()))
for _ in range(10):
    bench.gpu_step(call, ld, rd, "B", 8, up)
torch.cuda.synchronize()
# --- plain wall-clock of the three host entry points
acc = {}
def wrap(name):
    fn = getattr(ops, name)
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    setattr(ops, name, w)
for n in ("encode_launch", "encode_finish", "encode_backward"):
    wrap(n)
K = 200
t0 = time.perf_counter()
for _ in range(K):
    bench.gpu_step(call, ld, rd, "B", 8, up)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host loop {1e6*(t1-t0)/K:.0f} us/step (GPU may be the limiter)")
for n, v in acc.items():
    print(f"  {n:16s} {1e6*v/K:7.1f} us/step (2 calls)")
# --- profile
pr = cProfile.Profile()
pr.enable()
for _ in range(K):
    bench.gpu_step(call, ld, rd, "B", 8, up)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
