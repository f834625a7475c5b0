"""Float64 Gram V V^T (k = 1650, T = 20000) as one library GEMM vs the upper-triangular blocks of an nb x nb partition."""
import torch

k, T = 1650, 20000
torch.manual_seed(0)
v = torch.randn(k, T, device="cuda")


def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-44s %8.3f ms" % (name, e0.elapsed_time(e1) / reps))


def full():
    d = v.to(torch.float64)
    return d @ d.t()


def blocked(nb):
    d = v.to(torch.float64)
    edges = [round(i * k / nb / 8) * 8 for i in range(nb)] + [k]
    g = torch.empty((k, k), dtype=torch.float64, device="cuda")
    for i in range(nb):
        a = d[edges[i]:edges[i + 1]]
        for j in range(i, nb):
            blk = a @ d[edges[j]:edges[j + 1]].t()
            g[edges[i]:edges[i + 1], edges[j]:edges[j + 1]] = blk
            if j > i:
                g[edges[j]:edges[j + 1], edges[i]:edges[i + 1]] = blk.t()
    return g


timeit("to(float64) alone", lambda: v.to(torch.float64))
timeit("full d884gemm", full)
ref = full()
for nb in (2, 3, 4, 6):
    timeit("blocked nb=%d" % nb, lambda: blocked(nb))
    print("   max rel diff %.2e" % float((blocked(nb) - ref).abs().max() / ref.abs().max()))
