"""pmd_sym_product_f64 against the library float64 GEMM at the C2 sizes: final Gram (1650 x 20000 float32 rows) and the
whitening Gram M^T Z (11695 x 1650 float64 operands).  Usage: python scripts/debug/sym_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for (n, k) in [(1650, 20000), (1650, 11695), (3860, 30000)]:
    a = torch.randn((n, k), device=dev, generator=g)
    a64 = a.double()
    want = a64 @ a64.t()
    got = ops.sym_product_f64(a)
    err = float((got - want).abs().max() / want.abs().max())
    t_own = timeit(lambda: ops.sym_product_f64(a))
    t_lib = timeit(lambda: torch.matmul(a64, a64.t()))
    t_lib_cast = timeit(lambda: torch.matmul(a.double(), a.double().t()))
    nt = -(-n // 128)
    fl = nt * (nt + 1) // 2 * 128 * 128 * k * 2
    print("layout0 n=%d k=%d splits=%d: own %.3f ms (%.1f TFLOP/s on computed tiles), library full GEMM %.3f ms (with cast %.3f), "
          "max rel err %.1e" % (n, k, ops.sym_splits(n, k), t_own, fl / t_own / 1e9, t_lib, t_lib_cast, err), flush=True)
    m = a64.t().contiguous()           # (k, n) float64
    s = torch.randn((k,), device=dev, generator=g).double()
    z = m * s[:, None]
    want = m.t() @ z
    got = ops.sym_product_f64(m, z, layout=1)
    err = float((got - want).abs().max() / want.abs().max())
    t_own = timeit(lambda: ops.sym_product_f64(m, z, layout=1))
    t_lib = timeit(lambda: torch.matmul(m.t(), z))
    print("layout1 n=%d k=%d: own %.3f ms (%.1f TFLOP/s), library %.3f ms, max rel err %.1e" % (n, k, t_own, fl / t_own / 1e9, t_lib, err),
          flush=True)
    del a, a64, want, got, m, z
