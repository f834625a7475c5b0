"""Focused K1 benchmark: pmd_stats_pass_tc (tensor-core DFT) and pmd_stats_pass (SIMT FFT) on a synthetic movie.
Usage: python scripts/bench_k1.py [T] [reps] [tc|fft|both]   (512x512 FOV, float32)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from localmd_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = sys.argv[3] if len(sys.argv) > 3 else "both"
d = 512 * 512
dev = torch.device("cuda")
movie = torch.randn((T, d), device=dev) * 2 + 150
nbytes = 4.0 * d * T
res = {}
for name in (["tc", "fft"] if which == "both" else [which]):
    os.environ["PMD_K1"] = name
    ops.stats_pass(movie, T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        mp, npart, nv = ops.stats_pass(movie, T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[name] = (mp.sum(0), npart.sum(0) / nv)
    print("stats_pass[%s] %8.3f ms  %7.1f GB/s of movie bytes" % (name, ms, nbytes / ms / 1e6))
if len(res) == 2:
    print("mean  max rel diff %.2e" % ((res["tc"][0] / res["fft"][0] - 1).abs().max().item()))
    print("noise max rel diff %.2e" % ((res["tc"][1] / res["fft"][1] - 1).abs().max().item()))
