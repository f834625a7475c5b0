"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame / block partitioning, the ragged all-gather that
moves init frames and per-block results between ranks, frame gathering from owner ranks, and the frame-sharded
final Gram-SVD (all-reduce of the k x k Gram).  The CUDA kernels are not involved: these paths only move tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from localmd_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world=2, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_partitions():
    for T, w in [(20000, 8), (20000, 3), (1000, 4), (4096, 2), (5, 2)]:
        b = sharding.shard_bounds(T, w)
        assert b[0][0] == 0 and b[-1][1] == T and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(lo % 1024 == 0 for lo, _ in b if lo < T)
        assert sharding.owners_of([0, T - 1], b)[0] == [r for r, (lo, hi) in enumerate(b) if lo <= 0 < hi][0]
    p = sharding.block_partition(2601, 8)
    assert p[0][0] == 0 and p[-1][1] == 2601 and all(p[i][1] == p[i + 1][0] for i in range(7))
    assert max(hi - lo for lo, hi in p) - min(hi - lo for lo, hi in p) <= 1
    with pytest.raises(IndexError):
        sharding.owners_of([20000], sharding.shard_bounds(20000, 2))


def _ragged(rank, world):
    group = dist.group.WORLD
    for dtype in (torch.float32, torch.float64, torch.int32, torch.uint16, torch.uint8):
        counts = [3, 5] if world == 2 else [2] * world
        full = (torch.arange(sum(counts) * 7).reshape(sum(counts), 7) % 251).to(dtype)
        off = sum(counts[:rank])
        out = sharding.ragged_all_gather(full[off : off + counts[rank]].contiguous(), counts, group)
        assert out.dtype == dtype and torch.equal(out.to(torch.int64), full.to(torch.int64))
    # an empty contribution
    counts = [0, 4]
    full = torch.arange(8, dtype=torch.float32).reshape(4, 2)
    local = full[:0] if rank == 0 else full
    assert torch.equal(sharding.ragged_all_gather(local.contiguous(), counts, group), full)


def test_ragged_all_gather_gloo():
    _run(_ragged)


class _FakeMovie:
    """CPU stand-in for DeviceMovie: a frame shard of a (T, d) movie."""

    def __init__(self, full, lo, hi):
        self.full, self.lo, self.hi = full, lo, hi
        self.T_total, self.d = full.shape
        self.torch_dtype, self.device = full.dtype, full.device

    def gather(self, ids):
        assert all(self.lo <= i < self.hi for i in ids), "a rank may only read its own frames"
        return self.full[torch.as_tensor(list(ids), dtype=torch.int64)]


def _gather(rank, world):
    T, d = 5000, 6
    g = torch.Generator().manual_seed(0)
    for dtype in (torch.float32, torch.uint16):
        full = (torch.rand((T, d), generator=g) * 1000).to(dtype)
        bounds = sharding.shard_bounds(T, world)
        movie = _FakeMovie(full, *bounds[rank])
        ids = [4999, 0, 1023, 1024, 3000, 17, 4096, 2047]
        got = sharding.gather_frames(movie, ids, dist.group.WORLD, bounds)
        assert torch.equal(got.to(torch.float64), full[torch.tensor(ids)].to(torch.float64))
        run = list(range(2000, 2600))  # a contiguous init window straddling the shard boundary
        got = sharding.gather_frames(movie, run, dist.group.WORLD, bounds)
        assert torch.equal(got.to(torch.float64), full[2000:2600].to(torch.float64))


def test_gather_frames_gloo():
    _run(_gather)


def _exchange(rank, world):
    T, d1, d2 = 5000, 12, 5
    g = torch.Generator().manual_seed(1)
    starts = [(r, c) for r in (0, 2, 4, 6, 8) for c in (0, 1)]       # 10 blocks of height 4, row-major
    parts = sharding.block_partition(len(starts), world)
    ranges = sharding.block_row_ranges(starts, 4, parts)
    assert ranges[0][0] == 0 and ranges[-1][1] == d1 and all(ranges[i][1] >= ranges[i + 1][0] for i in range(world - 1))
    owned = sharding.owned_row_ranges(ranges, d1)
    assert owned[0][0] == 0 and owned[-1][1] == d1 and all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))
    for dtype in (torch.float32, torch.uint16):
        full = (torch.rand((T, d1 * d2), generator=g) * 1000).to(dtype)
        bounds = sharding.shard_bounds(T, world)
        movie = _FakeMovie(full, *bounds[rank])
        for ids in (list(range(2000, 2600)), [4999, 0, 1023, 1024, 3000, 17], list(range(100, 140))):
            got = sharding.exchange_frame_rows(movie, ids, ranges, d2, dist.group.WORLD, bounds)
            lo, hi = ranges[rank]
            want = full[torch.tensor(ids)][:, lo * d2 : hi * d2]
            assert got.shape == want.shape and torch.equal(got.to(torch.float64), want.to(torch.float64))


def test_exchange_frame_rows_gloo():
    _run(_exchange)
    _run(_exchange, 3)


def _svd(rank, world):
    from localmd_b200 import ops
    from localmd_b200.decomposition import projected_svd

    # the product has no CPU path: this test covers the COLLECTIVE logic of the frame-sharded final SVD (Gram all-reduce,
    # local Vt block), so the two device products are replaced by plain torch stand-ins for the duration of the test
    ops.sym_product_f64 = lambda a, b=None, layout=0: a.double() @ a.double().t()
    ops.matmul_3xtf32_any = lambda a, b: a @ b

    rng = np.random.default_rng(3)
    k, T, R = 130, 3000, 150
    data = (rng.standard_normal((k, 40)) @ rng.standard_normal((40, T)) + 0.1 * rng.standard_normal((k, T))).astype(np.float32)
    proj = rng.standard_normal((R, k)).astype(np.float32)
    bounds = sharding.shard_bounds(T, world)
    lo, hi = bounds[rank]
    r_sh, s_sh, vt_sh = projected_svd(torch.from_numpy(proj), torch.from_numpy(data[:, lo:hi].copy()), dist.group.WORLD)
    r_full, s_full, vt_full = projected_svd(torch.from_numpy(proj), torch.from_numpy(data))
    np.testing.assert_allclose(s_sh.numpy(), s_full.numpy(), rtol=1e-5)
    # the factorisation itself (signs of singular vectors are free): R diag(s) Vt restricted to the shard
    a = (r_sh.numpy() * s_sh.numpy()[None]) @ vt_sh.numpy()
    b = (r_full.numpy() * s_full.numpy()[None]) @ vt_full.numpy()[:, lo:hi]
    assert np.abs(a - b).max() < 1e-3 * np.abs(b).max()
    vt = sharding.ragged_all_gather(vt_sh.t().contiguous(), [h - l for l, h in bounds], dist.group.WORLD).t()
    np.testing.assert_allclose(np.abs((vt @ vt.t()).numpy() - np.eye(k)).max(), 0, atol=1e-3)
    # fewer frames than components (k > T): the column blocks are exchanged and every rank solves the T x T problem
    T2 = 120     # (above 112, where the small-matrix Jacobi kernel -- device only -- would take over)
    data2 = rng.standard_normal((k, T2)).astype(np.float32)
    lo2, hi2 = [(0, 47), (47, T2)][rank] if world == 2 else sharding.shard_bounds(T2, world)[rank]
    r2s, s2s, vt2s = projected_svd(torch.from_numpy(proj), torch.from_numpy(data2[:, lo2:hi2].copy()), dist.group.WORLD)
    r2f, s2f, vt2f = projected_svd(torch.from_numpy(proj), torch.from_numpy(data2))
    assert tuple(vt2s.shape) == (T2, hi2 - lo2)
    np.testing.assert_allclose(s2s.numpy(), s2f.numpy(), rtol=1e-6)
    np.testing.assert_allclose(vt2s.numpy(), vt2f.numpy()[:, lo2:hi2], atol=1e-6)
    np.testing.assert_allclose(r2s.numpy(), r2f.numpy(), rtol=1e-5, atol=1e-5)


def test_sharded_projected_svd_gloo():
    _run(_svd)
