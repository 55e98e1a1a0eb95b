"""Host vs device timeline of one bench step (mode B): where does the step wait?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from radardistill_b200 import vfe as V
lidar, radar = bench.make_clouds(0, 8)
dev = torch.device("cuda", 0)
lid, rad, call = bench.build_modules(dev, "B", False)
ld, rd = torch.from_numpy(lidar).to(dev), torch.from_numpy(radar).to(dev)
up = bench.make_upstream(dev, len(lidar), len(radar), list(lid.parameters()) + list(rad.parameters()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream()
side = V._side_stream(dev)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(main); return e
for it in range(8):
    flush.fill_(1)
    for p_ in up['params']: p_.grad = None
    torch.cuda.synchronize()
    H = [time.perf_counter()]; E = [ev()]
    bd = {"points": ld, "radar_points": rd, "batch_size": 8}
    side.wait_stream(main)
    tok1 = lid.launch(bd); H.append(time.perf_counter()); E.append(ev())
    with torch.cuda.stream(side):
        tok2 = rad.launch(bd)
    H.append(time.perf_counter()); E.append(ev())
    bd = lid.finish(bd, tok1); H.append(time.perf_counter()); E.append(ev())
    with torch.cuda.stream(side):
        bd = rad.finish(bd, tok2)
    main.wait_stream(side)
    H.append(time.perf_counter()); E.append(ev())
    outs = [bd["radar_pillar_features"], bd["pillar_features"]]
    gr = [up["radar"][:outs[0].shape[0]], up["lidar"][:outs[1].shape[0]]]
    torch.autograd.backward(outs, gr); H.append(time.perf_counter()); E.append(ev())
    torch.cuda.synchronize(); H.append(time.perf_counter())
    names = ["lidar launch", "radar launch", "lidar finish", "radar finish+join", "backward"]
    print(f"--- step {it}: total device {E[0].elapsed_time(E[-1])*1e3:.0f} us, wall {1e6*(H[-1]-H[0]):.0f} us")
    for i, n in enumerate(names):
        print(f"   {n:20s} host +{1e6*(H[i+1]-H[i]):7.0f} us   device +{E[i].elapsed_time(E[i+1])*1e3:7.0f} us")
