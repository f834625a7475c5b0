"""Timing of symmetric eigensolvers for the final k x k Gram (k = 1650)."""
import torch

k = 1650
torch.manual_seed(0)
v = torch.randn(k, 20000, device="cuda", dtype=torch.float64) * torch.logspace(0, -3, k, device="cuda", dtype=torch.float64)[:, None]
g = v @ v.t()


def timeit(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-36s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


timeit("eigh f64", lambda: torch.linalg.eigh(g))
g32 = g.float()
timeit("eigh f32", lambda: torch.linalg.eigh(g32))
timeit("eigvalsh f64", lambda: torch.linalg.eigvalsh(g))
timeit("matmul f64 k^3", lambda: g @ g)
timeit("matmul f32 k^3", lambda: g32 @ g32)
timeit("cholesky f64", lambda: torch.linalg.cholesky(g + 1e-3 * torch.eye(k, device="cuda", dtype=torch.float64)))
timeit("qr f64", lambda: torch.linalg.qr(g))
for lib in ("cusolver", "magma"):
    try:
        torch.backends.cuda.preferred_linalg_library(lib)
        timeit("eigh f64 (%s)" % lib, lambda: torch.linalg.eigh(g))
    except Exception as exc:
        print(lib, "unavailable:", exc)
torch.backends.cuda.preferred_linalg_library("cusolver")
for n in (412, 825, 1100):
    gs = g[:n, :n].contiguous()
    timeit("eigh f64 n=%d" % n, lambda: torch.linalg.eigh(gs))
gs = [g[i * 412:(i + 1) * 412, i * 412:(i + 1) * 412].contiguous() for i in range(4)]
ss = [torch.cuda.Stream() for _ in range(4)]
def four():
    cur = torch.cuda.current_stream()
    for s, m in zip(ss, gs):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            torch.linalg.eigh(m)
    for s in ss:
        cur.wait_stream(s)
timeit("4 x eigh f64 n=412 on 4 streams", four)
gb = torch.stack(gs)
timeit("batched eigh f64 4 x 412", lambda: torch.linalg.eigh(gb))
