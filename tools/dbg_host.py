import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
lidar, radar = bench.make_clouds(0, 8)
dev = torch.device("cuda", 0)
lid, rad, call = bench.build_modules(dev, "A", False)
ld, rd = torch.from_numpy(lidar).to(dev), torch.from_numpy(radar).to(dev)
bd = {"points": ld, "radar_points": rd, "batch_size": 8}
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tok = rad.launch(bd); t1 = time.perf_counter()
    rad.finish(bd, tok); t2 = time.perf_counter()
    loss = bd["radar_pillar_features"].sum(); t3 = time.perf_counter()
    loss.backward(); t4 = time.perf_counter(); torch.cuda.synchronize(); t5 = time.perf_counter()
    with torch.no_grad():
        tl = lid.launch(bd); t6 = time.perf_counter(); lid.finish(bd, tl); t7 = time.perf_counter()
    print(f"radar launch {1e6*(t1-t0):6.0f} us finish {1e6*(t2-t1):6.0f} sum {1e6*(t3-t2):5.0f} backward(host) {1e6*(t4-t3):6.0f} drain {1e6*(t5-t4):5.0f} | lidar launch {1e6*(t6-t5):6.0f} finish {1e6*(t7-t6):6.0f}")
