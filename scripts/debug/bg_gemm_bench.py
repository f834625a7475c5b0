"""Timing of the skinny GEMMs of the background rSVD (pmd_loader.py:57-62) at C2 size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops

d, n, l = 512 * 512, 1000, 25
a_t = torch.randn(n, d, device="cuda")
sk = torch.randn(n, l, device="cuda")
q = torch.linalg.qr(torch.randn(d, l, device="cuda"))[0].contiguous()


def timeit(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


timeit("y = a_t.t() @ sk            (d x n)(n x l)", lambda: torch.matmul(a_t.t(), sk))
timeit("bmat = q.t() @ a_t.t()      (l x d)(d x n)", lambda: torch.matmul(q.t(), a_t.t()))
timeit("bmat^T = a_t @ q            (n x d)(d x l)", lambda: torch.matmul(a_t, q))
timeit("3xTF32 a_t @ q", lambda: ops.matmul_3xtf32_any(a_t, q))
timeit("3xTF32 a_t.t() @ sk", lambda: ops.matmul_3xtf32_any(a_t.t().contiguous(), sk))
timeit("u = q @ e                   (d x l)(l x 15)", lambda: torch.matmul(q, sk[:l, :15]))
