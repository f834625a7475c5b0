import time, torch, numpy as np
x = torch.empty((4096, 512*512), dtype=torch.float32, pin_memory=True)
x.fill_(1.0)
y = torch.from_numpy(x.numpy())
print("from_numpy is_pinned:", y.is_pinned(), " view is_pinned:", y.view(4096, -1)[100:200].is_pinned())
d = torch.empty_like(x, device="cuda")
for name, src in [("pinned tensor", x), ("from_numpy view", y)]:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d.copy_(src, non_blocking=True); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(name, "%.1f GB/s" % (x.numel() * 4 / dt / 1e9))
s = torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s):
    for i in range(0, 4096, 512):
        d[i:i+512].copy_(y[i:i+512], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("chunked side stream %.1f GB/s" % (x.numel() * 4 / dt / 1e9))
