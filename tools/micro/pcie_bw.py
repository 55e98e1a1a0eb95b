import torch, time
dev = torch.device("cuda", 0)
n = 191 << 20
d = torch.empty(n, dtype=torch.uint8, device=dev); h = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(66 << 20, dtype=torch.uint8, device=dev); h2 = torch.empty(66 << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
a = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 191MiB: {a*1e3:.2f} ms  {n/a/1e9:.1f} GB/s")
b = t(lambda: d2.copy_(h2, non_blocking=True)); print(f"H2D 66MiB: {b*1e3:.2f} ms  {(66<<20)/b/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
c = t(both); print(f"both concurrently: {c*1e3:.2f} ms")
