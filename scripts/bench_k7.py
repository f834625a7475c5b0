"""Focused K7 benchmark: project_supertile / project_local / project_dense on a synthetic block-structured U.
Usage: python scripts/bench_k7.py [T] [reps]   (512x512 FOV, 20x20 blocks, ranks ~ 1+Poisson(3.5))"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from localmd_b200 import ops  # noqa: E402
from localmd_b200.decomposition import tile_starts  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = sys.argv[3] if len(sys.argv) > 3 else "all"
d1 = int(os.environ.get("PMD_BK7_D1", "512"))
d2 = 512
bh = bw = 20
K = 15
rng = np.random.default_rng(0)
rows, cols = tile_starts(d1, bh), tile_starts(d2, bw)
nb = len(rows) * len(cols)
ranks = (1 + rng.poisson(3.5, nb)).clip(1, 13).astype(np.int32)
col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]]).astype(np.int64)
n_local = int(ranks.sum())
dev = torch.device("cuda")
uv = torch.randn((n_local, bh * bw), device=dev)
bg = torch.randn((K, d1 * d2), device=dev)
PAD = int(os.environ.get("PMD_BK7_PAD", "0"))  # extra floats per frame (frame stride experiment)
movie = torch.randn((T, d1 * d2 + PAD), device=dev) * 3 + 100
mean = torch.full((d1 * d2,), 100.0, device=dev)
inv = torch.full((d1 * d2,), 0.5, device=dev)
starts = torch.from_numpy(np.array([(r, c) for r in rows for c in cols], dtype=np.int32)).to(dev)
st = ops.make_supertiles(rows, cols, bh, bw, ranks, col0)
print("G", st["G"], "tiles", len(st["tiles"]), "tasks", len(st["tasks"]), "max tasks/tile", int(np.diff(st["task_ptr"]).max()),
      "mean", float(np.diff(st["task_ptr"]).mean()), "region", st["max_h"], st["max_w"], "mean rank", ranks.mean())
std = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
ranks_d, col0_d = torch.from_numpy(ranks).to(dev), torch.from_numpy(col0).to(dev)
tasks = torch.from_numpy(ops.make_tasks(ranks)).to(dev)
z = torch.zeros((n_local + K, T), device=dev)
nbytes = 4.0 * d1 * d2 * T


def timeit(name, fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-18s %8.3f ms  %7.1f GB/s of movie bytes  (x%.1f to T=20000: %.1f ms)" % (name, ms, nbytes / ms / 1e6, 20000 / T, ms * 20000 / T))


if which in ("all", "stream"):
    sst = ops.make_strips(rows, cols, bh, bw, d1, d2, ranks, col0, K)
    sst_d = {k: (torch.from_numpy(v).to(dev) if k in ("items", "slot_ptr", "tasks") else v) for k, v in sst.items()}
    upack = ops.pack_strip_u(sst, uv, bg, bh * bw)
    print("strips: items", sst["n_items"], "max_rw", sst["max_rw"], "tasks", len(sst["tasks"]))
    timeit("project_stream", lambda: ops.project_stream(movie, d2, sst_d, upack, mean, inv, z[:n_local], z[n_local:]))
if which in ("all", "tc", "stream"):
    G = int(os.environ.get("PMD_TC_G", "0")) or None
    tst = ops.make_strips_tc(rows, cols, bh, bw, d1, d2, ranks, col0, K, G=G)
    tst_d = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in tst.items()}
    bimg = ops.pack_strips_tc(tst_d, uv, bg, bh * bw, d2)
    it = tst["items"]
    print("tc strips: G", tst["G"], "items", tst["n_items"], "max w8", tst["max_w8"], "image MB", bimg.numel() / 1e6,
          "streamed/ideal %.3f" % (float((it[:, 1] * it[:, 3]).sum()) * 8 / (d1 * d2)))
    z2 = torch.zeros_like(z)
    timeit("pack_strips_tc", lambda: ops.pack_strips_tc(tst_d, uv, bg, bh * bw, d2))
    timeit("project_stream_tc", lambda: ops.project_stream_tc(movie, d2, tst_d, bimg, mean, inv, z2[:n_local], z2[n_local:]))
    if which in ("all", "stream"):
        err = (z2 - z).abs().max().item() / z.abs().max().item()
        print("max |tc - simt| / max |simt| = %.3e" % err)
    n = min(T, 512)
    ref = torch.zeros((n_local + K, n), dtype=torch.float64, device=dev)
    yc = ((movie[:n].double() - mean.double()) * inv.double())
    ref[n_local:] = bg.double() @ yc.t()
    qi, qj = np.divmod(np.arange(bh * bw), bw)
    for b in range(0, nb, max(1, nb // 64)):
        i0, j0 = rows[b // len(cols)], cols[b % len(cols)]
        pix = torch.from_numpy((i0 + qi) * d2 + j0 + qj).to(dev)
        ref[col0[b] : col0[b] + ranks[b]] = uv[col0[b] : col0[b] + ranks[b]].double() @ yc[:, pix].t()
    sel = ref.abs().sum(1) > 0
    err = ((z2[:, :n].double() - ref)[sel].abs().max() / ref.abs().max()).item()
    print("tc vs float64 reference (sampled blocks + background): max abs err / max |ref| = %.3e" % err)
if which in ("all", "ts"):
    Wf = int(os.environ.get("PMD_TS_W", "0")) or None
    Nf = int(os.environ.get("PMD_TS_N", "0")) or None
    sst = ops.make_strips_ts(rows, cols, bh, bw, d1, d2, ranks, col0, K, W=Wf, N=Nf)
    sst_d = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in sst.items()}
    bimg_ts = ops.pack_strips_ts(sst_d, uv, bg, inv, bh * bw, d2)
    it = sst["items"]
    print("ts strips: W", sst["W"], "N", sst["N"], "tiles", sst["tiles"], "items", sst["n_items"], "image MB", bimg_ts.numel() / 1e6,
          "streamed/ideal %.3f" % (float((it[:, 1] * it[:, 3]).sum()) * 32 / (d1 * d2)))
    z3 = torch.zeros_like(z)
    timeit("pack_strips_ts", lambda: ops.pack_strips_ts(sst_d, uv, bg, inv, bh * bw, d2))
    timeit("project_stream_ts", lambda: ops.project_stream_ts(movie, d2, sst_d, bimg_ts, mean, z3[:n_local], z3[n_local:]))
    n = min(T, 512)
    ref = torch.zeros((n_local + K, n), dtype=torch.float64, device=dev)
    yc = ((movie[:n, : d1 * d2].double() - mean.double()) * inv.double())
    ref[n_local:] = bg.double() @ yc.t()
    qi, qj = np.divmod(np.arange(bh * bw), bw)
    for b in range(0, nb, max(1, nb // 64)):
        i0, j0 = rows[b // len(cols)], cols[b % len(cols)]
        pix = torch.from_numpy((i0 + qi) * d2 + j0 + qj).to(dev)
        ref[col0[b] : col0[b] + ranks[b]] = uv[col0[b] : col0[b] + ranks[b]].double() @ yc[:, pix].t()
    sel = ref.abs().sum(1) > 0
    err = ((z3[:, :n].double() - ref)[sel].abs().max() / ref.abs().max()).item()
    print("ts vs float64 reference (sampled blocks + background): max abs err / max |ref| = %.3e" % err)
if which in ("all", "supertile"):
    timeit("project_supertile", lambda: ops.project_supertile(movie, d2, std, bh, bw, uv, mean, inv, z[:n_local]))
if which in ("all", "local"):
    timeit("project_local(v1)", lambda: ops.project_local(movie, d2, starts, bh, bw, ranks_d, col0_d, tasks, uv, mean, inv, z[:n_local]))
if which in ("all", "dense"):
    timeit("project_dense", lambda: ops.project_dense(movie, bg, mean, inv, z[n_local:]))
