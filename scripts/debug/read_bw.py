"""Calibration: read-only bandwidth of this GPU for sequential and strip-shaped (frame-strided) access."""
import torch

T, d1, d2 = 20000, 512, 512
x = torch.empty((T, d1, d2), device="cuda").normal_()


def timeit(name, fn, nbytes, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-40s %8.3f ms  %7.1f GB/s" % (name, ms, nbytes / ms / 1e6))


n = x.numel() * 4
timeit("sum() sequential", lambda: x.sum(), n)
timeit("sum(dim=0) (per-pixel over frames)", lambda: x.sum(dim=0), n)
timeit("sum(dim=(1,2)) (per-frame)", lambda: x.sum(dim=(1, 2)), n)
for w in (32, 96, 128, 256, 512):
    timeit("strip [:, :, 64:64+%d].sum(dim=0)" % w, lambda: x[:, :, 64 : 64 + w].sum(dim=0), T * d1 * w * 4)
y = torch.empty_like(x)
timeit("copy_ (read+write)", lambda: y.copy_(x), 2 * n)
