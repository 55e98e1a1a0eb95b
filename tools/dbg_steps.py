import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
lidar, radar = bench.make_clouds(0, 8)
dev = torch.device("cuda", 0)
for mode in ("B", "A"):
    lid, rad, call = bench.build_modules(dev, mode, False)
    ld, rd = torch.from_numpy(lidar).to(dev), torch.from_numpy(radar).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for i in range(14):
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        bench.gpu_step(call, ld, rd, mode, 8)
        e.record()
        torch.cuda.synchronize()
        print(mode, i, f"ev={s.elapsed_time(e):8.3f} ms wall={(time.perf_counter()-t0)*1e3:8.3f} ms alloc={torch.cuda.memory_allocated()/1e9:.2f} GB reserved={torch.cuda.memory_reserved()/1e9:.2f} GB", flush=True)
