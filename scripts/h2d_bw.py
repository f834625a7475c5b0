"""Raw pinned host -> device copy bandwidth of this box (context for the e2e number of bench.py)."""
import time
import torch

n = 2 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("pinned H2D: %.1f GB/s" % (4 * n / dt / 1e9))
t0 = time.perf_counter()
for _ in range(4):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("pinned D2H: %.1f GB/s" % (4 * n / dt / 1e9))
