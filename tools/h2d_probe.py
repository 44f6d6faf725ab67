"""Scratch: pinned host -> device copy bandwidth on this box (one stream, large and 4.7 MB copies; two streams)."""
import time
import torch

dev = torch.device("cuda", 0)
for mb in (4.72, 75.5, 1024):
    n = int(mb * 1e6) // 4
    h = torch.empty(n, dtype=torch.int32).pin_memory()
    d = torch.empty(n, dtype=torch.int32, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    reps = max(3, int(2000 / mb))
    t = time.perf_counter()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"H2D {mb:8.2f} MB x{reps}: {reps * n * 4 / dt / 1e9:6.1f} GB/s")
    t = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"D2H {mb:8.2f} MB x{reps}: {reps * n * 4 / dt / 1e9:6.1f} GB/s")
n = int(75.5e6) // 4
hs = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
ds = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(2)]
ss = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    for i in range(2):
        with torch.cuda.stream(ss[i]):
            ds[i].copy_(hs[i], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print(f"H2D two streams: {40 * n * 4 / dt / 1e9:6.1f} GB/s")
