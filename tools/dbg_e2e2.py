import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from radardistill_b200 import pipeline
lidar, radar = bench.make_clouds(0, 8)
dev = torch.device("cuda", 0)
mode = "B"
lid, rad, call = bench.build_modules(dev, mode, False)
host_in = {"points": torch.from_numpy(lidar).pin_memory(), "radar_points": torch.from_numpy(radar).pin_memory()}
T = {}
def step(d):
    t0 = time.perf_counter()
    torch.cuda.current_stream().synchronize()
    T["wait_up"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = bench.gpu_step(call, d["points"], d["radar_points"], mode, 8)
    torch.cuda.current_stream().synchronize()
    T["compute"] = time.perf_counter() - t0
    return r
pipe = pipeline.HostPipeline(step, dev, ("pillar_features", "pillar_coords", "radar_pillar_features", "radar_pillar_coords"))
for i in range(12):
    t0 = time.perf_counter()
    pipe.submit(host_in, host_in)
    t1 = time.perf_counter()
    print(i, f"submit {1e3*(t1-t0):7.3f} ms wait_up {1e3*T['wait_up']:.3f} compute {1e3*T['compute']:.3f}", flush=True)
