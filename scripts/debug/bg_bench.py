"""Timing of the background-basis kernels at C2 size (1000 random frames of a 512x512x20000 float32 movie)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from localmd_b200 import ops

d, T, n, l = 512 * 512, 6000, 1000, 25
torch.manual_seed(0)
movie = torch.randn((T, d), device="cuda")
frames = torch.from_numpy(np.sort(np.random.default_rng(0).choice(T, n, replace=False))).cuda()
mean = torch.zeros(d, device="cuda"); std = torch.ones(d, device="cuda")
om = torch.randn((n, l), device="cuda")


def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-40s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


for ld in (1000, 1024):
    timeit("standardize_frames_t ld=%d" % ld, lambda: ops.standardize_frames_t(movie, frames, mean, std, ld=ld))
yt = ops.standardize_frames_t(movie, frames, mean, std, ld=1024)
timeit("rows_sketch l=25", lambda: ops.rows_sketch(yt, n, om))
timeit("rows_sketch l=13", lambda: ops.rows_sketch(yt, n, om[:, :13].contiguous()))
timeit("torch yt[:, :n] @ om", lambda: yt[:, :n] @ om)
y = ops.rows_sketch(yt, n, om)
ref = yt[:, :n].double() @ om.double()
print("sketch rel err %.2e" % float((y.double() - ref).abs().max() / ref.abs().max()))
timeit("gram_cols", lambda: ops.gram_cols(y[None], l))
g = ops.gram_cols(y[None], l)
timeit("chol_whiten", lambda: ops.chol_whiten(g))
t = ops.chol_whiten(g)
timeit("rows_times_small", lambda: ops.rows_times_small(y[None], t))
q = ops.rows_times_small(y[None], t, transposed=True)[0]
for nr in (148, 296, 592):
    timeit("bg_project_t k=25 ranges=%d" % nr, lambda: ops.bg_project_t(yt, q, n_ranges=nr))
timeit("torch q @ yt", lambda: q @ yt)
