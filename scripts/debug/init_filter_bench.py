"""Timing of the init-filter pieces at C2 size (5000 init frames, 512x512)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops

d, t, K, T = 512 * 512, 5000, 15, 6000
dev = torch.device("cuda")
movie = torch.randn((T, d), device=dev)
mean = torch.zeros(d, device=dev)
std = torch.ones(d, device=dev)
idx = torch.arange(500, 500 + t, device=dev)
bg = torch.linalg.qr(torch.randn((d, K), device=dev))[0].t().contiguous()


def timeit(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-28s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


yt = ops.standardize_frames_t(movie, idx, mean, std)
print(yt.shape)
timeit("standardize_frames_t", lambda: ops.standardize_frames_t(movie, idx, mean, std))
timeit("torch bg @ yt", lambda: torch.matmul(bg, yt))
vbg = torch.matmul(bg, yt).contiguous()
timeit("torch addmm_", lambda: yt.addmm_(bg.t(), vbg, alpha=-1.0))
timeit("bg_project_t", lambda: ops.bg_project_t(yt, bg))
timeit("bg_remove_t", lambda: ops.bg_remove_t(yt, bg, vbg))
