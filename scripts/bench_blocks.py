"""Block-stage contraction micro-benchmark at C2 (2601 blocks of 20x20, t=5000, r=50): the generations of the block
projection (tcgen05 with the movie operand staged in shared memory / in tensor memory) and the spatial projection.
Usage: python scripts/bench_blocks.py [project|spatial|all]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from localmd_b200 import ops
from localmd_b200.decomposition import tile_starts
which = sys.argv[1] if len(sys.argv) > 1 else "all"
d1 = d2 = 512; bh = bw = 20; t = 5000; r = 50; rp = 52
dev = torch.device("cuda")
rows, cols = tile_starts(d1, bh), tile_starts(d2, bw)
starts = torch.tensor([(a, c) for a in rows for c in cols], dtype=torch.int32, device=dev)
nb = starts.shape[0]
yt = torch.randn((d1 * d2, t), device=dev)
w = torch.zeros((nb, bh * bw, rp), device=dev); w[:, :, :r] = torch.randn((nb, bh * bw, r), device=dev)
def timeit(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * nb * bh * bw * r * t
    print("%-28s %7.2f ms  %6.1f TFLOP/s (useful fp32-equivalent)" % (name, ms, fl / ms / 1e9))
    return out
if which in ("all", "project"):
    os.environ["PMD_BLOCK_PROJECT"] = "tc"
    b = timeit("block_project_tc (smem A)", lambda: ops.block_project_tc(yt, 0, t, d2, starts, bh, bw, w, r))
    os.environ["PMD_BLOCK_PROJECT"] = "ts"
    c = timeit("block_project_ts (tmem A)", lambda: ops.block_project_tc(yt, 0, t, d2, starts, bh, bw, w, r))
    print("max |ts - tc| / max|tc| = %.2e" % ((c - b).abs().max() / b.abs().max()).item())
if which in ("all", "spatial"):
    vb = torch.randn((nb, r, t), device=dev)
    os.environ["PMD_BLOCK_SPATIAL"] = "tc"
    d = timeit("block_spatial_tc (smem A)", lambda: ops.block_spatial_tc(yt, 0, t, d2, starts, bh, bw, vb, rp))
    os.environ["PMD_BLOCK_SPATIAL"] = "ts"
    e = timeit("block_spatial_ts (tmem A)", lambda: ops.block_spatial_tc(yt, 0, t, d2, starts, bh, bw, vb, rp))
    print("max |ts - tc| / max|tc| = %.2e" % ((e - d).abs().max() / d.abs().max()).item())
