"""Rotation of the block temporal components into singular vectors, V <- L^T V (decomposition.py:319-323), through the
tensor-core block projection kernel (V[b] seen as a 1 x r pixel block of a pixel-major movie) against the library batched
GEMM.  Usage: python scripts/debug/rot_bench.py [nb] [r] [ld]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2601
r = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ld = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
rp = (r + 3) // 4 * 4
vn = torch.randn((nb, r, ld), device=dev, generator=g)
lmat = torch.linalg.qr(torch.randn((nb, r, r), device=dev, generator=g))[0].contiguous()
lpad = torch.zeros((nb, r, rp), device=dev)
lpad[:, :, :r] = lmat
zero = torch.zeros((nb, 2), dtype=torch.int32, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def lib():
    with ops.fp32_matmul():
        return torch.bmm(lmat.transpose(1, 2), vn)


def own():
    return ops.block_project_tc(vn, r * ld, ld, r, zero, 1, r, lpad, r)


want = torch.bmm(lmat[:64].double().transpose(1, 2), vn[:64].double())
e_lib = float((lib()[:64].double() - want).abs().max() / want.abs().max())
e_own = float((own()[:64].double() - want).abs().max() / want.abs().max())
print("nb %d r %d ld %d: library %.3f ms (err %.1e), block_project_ts %.3f ms (err %.1e); bytes %.2f GB" % (
    nb, r, ld, timeit(lib), e_lib, timeit(own), e_own, 2 * vn.numel() * 4 / 1e9), flush=True)
