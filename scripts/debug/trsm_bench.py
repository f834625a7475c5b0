"""Whitening solve P = M L^-T (R = 11695, k = 1650, float64): library triangular solve vs explicit inverse + GEMM."""
import torch
R, k = 11696, 1650
torch.manual_seed(0)
m = torch.randn(R, k, device="cuda", dtype=torch.float64)
g = m.t() @ m
l = torch.linalg.cholesky(g)


def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-44s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


eye = torch.eye(k, device="cuda", dtype=torch.float64)
timeit("cholesky_ex", lambda: torch.linalg.cholesky_ex(g))
timeit("solve_triangular(L, M^T)", lambda: torch.linalg.solve_triangular(l, m.t(), upper=False))
timeit("solve_triangular(L, I)", lambda: torch.linalg.solve_triangular(l, eye, upper=False))
timeit("solve_triangular(L^T, M, left=False)", lambda: torch.linalg.solve_triangular(l.t(), m, upper=True, left=False))
linv = torch.linalg.solve_triangular(l, eye, upper=False)
timeit("M @ Linv^T", lambda: m @ linv.t())
a = torch.linalg.solve_triangular(l, m.t(), upper=False).t()
b = m @ linv.t()
print("rel diff %.2e" % float((a - b).abs().max() / a.abs().max()))
timeit("cholesky_inverse-free: inv via chol_solve", lambda: torch.cholesky_solve(eye, l))
