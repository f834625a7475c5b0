"""Timing of the batched Jacobi eigensolver at the block-stage sizes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops

torch.manual_seed(0)
nb = 2601
for n, m, f32 in ((50, 5000, False), (60, 500, True)):
    x = torch.randn(nb, n, m, device="cuda") * torch.logspace(0, -2, n, device="cuda")[None, :, None]
    g = ops.gram_rows(x)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc = g.clone()
        torch.cuda.synchronize()
        e0.record()
        w, v = ops.jacobi_eigh(gc, mode=0, sweeps_f32=f32)
        e1.record()
        torch.cuda.synchronize()
    print("jacobi n=%d f32=%s: %.3f ms" % (n, f32, e0.elapsed_time(e1)))
for (m, n, ld) in ((400, 50, 52), (100, 60, 60)):
    s = torch.randn(nb, m, ld, device="cuda")
    s[:, :, n:] = 0
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sc = s.clone()
        torch.cuda.synchronize()
        e0.record()
        ops.block_orth(sc, n)
        e1.record()
        torch.cuda.synchronize()
    print("block_orth m=%d n=%d: %.3f ms" % (m, n, e0.elapsed_time(e1)))
x = torch.randn(nb, 50, 5000, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    ops.gram_rows(x)
    e1.record()
    torch.cuda.synchronize()
print("gram_rows n=50 m=5000: %.3f ms" % e0.elapsed_time(e1))
