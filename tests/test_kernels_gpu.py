"""Kernel-level parity tests: every C-ABI entry point (through localmd_b200.ops -> ctypes ->
libpmd_sm100.so) against the CPU oracle / plain NumPy float64 on seeded inputs.  Tolerances are
written next to each comparison (float32 arithmetic with reordered sums unless stated)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import oracle.pmd_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from localmd_b200 import ops as _ops

    return _ops


def dev(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize(
    "T,d1,d2,dtype",
    [(1300, 9, 13, np.float32), (2048 + 544, 8, 16, np.float32), (1024 + 176, 5, 7, np.uint16), (300, 6, 6, np.int16),
     (200, 4, 5, np.float32), (700, 3, 70, np.uint8), (1100, 4, 4, np.float64), (2300, 4, 4, np.int32)],
)
def test_stats_pass(ops, T, d1, d2, dtype):
    rng = np.random.default_rng(T + d1)
    base = rng.uniform(50, 200, size=(d1, d2))
    sig = rng.uniform(0.5, 4, size=(d1, d2))
    y = base[None] + sig[None] * rng.standard_normal((T, d1, d2))
    if np.issubdtype(dtype, np.integer):
        y = np.clip(np.rint(y), 0, np.iinfo(dtype).max)
    movie = y.astype(dtype)
    mean_ref, std_ref = O.mean_and_noise(movie)
    mp, npart, n_var = ops.stats_pass(dev(movie).view(T, d1 * d2), T)
    mean = mp.sum(0).cpu().numpy().reshape(d1, d2)
    np.testing.assert_allclose(mean, mean_ref, rtol=2e-6)  # float32 rounding of chunk sums
    if T >= 256:
        std = (npart.sum(0) / n_var).cpu().numpy().reshape(d1, d2)
        np.testing.assert_allclose(std, std_ref, rtol=2e-5)  # fp32 DFT contraction vs scipy's float32 FFT
    else:
        assert n_var == 0 and float(npart.abs().max()) == 0.0


@pytest.mark.parametrize(
    "T,d1,d2,dtype",
    [(1024, 8, 16, np.float32), (2048 + 544, 13, 12, np.float32), (700, 16, 24, np.uint16), (256, 4, 4, np.float32),
     (300, 16, 9, np.float32), (1500, 16, 16, np.int16), (100, 8, 8, np.float32), (3000, 40, 36, np.uint8),
     (1300, 12, 10, np.float64), (1024 + 255, 20, 20, np.int32), (4096 + 130, 64, 66, np.float32)],
)
@pytest.mark.parametrize("which", ["tc", "fft"])
def test_stats_pass_both_kernels(ops, T, d1, d2, dtype, which, monkeypatch):
    """Both generations of K1 (tensor-core DFT behind 2-D TMA, and the SIMT FFT kernel the unaligned movies fall back to)
    against the oracle's mean / Welch estimate; includes a movie with strong slow signals (rounding of the DFT matrix)."""
    monkeypatch.setenv("PMD_K1", which)
    rng = np.random.default_rng(T + d1)
    base = rng.uniform(50, 200, size=(d1, d2))
    sig = rng.uniform(0.5, 4, size=(d1, d2))
    tt = np.arange(T)
    slow = 40.0 * np.exp(-((tt - 100) % 300) / 15.0)[:, None, None] * rng.uniform(0, 1, size=(1, d1, d2)) + 0.01 * tt[:, None, None]
    y = base[None] + sig[None] * rng.standard_normal((T, d1, d2)) + slow
    if np.issubdtype(dtype, np.integer):
        y = np.clip(np.rint(y), 0, np.iinfo(dtype).max)
    movie = y.astype(dtype)
    mean_ref, std_ref = O.mean_and_noise(movie)
    mp, npart, n_var = ops.stats_pass(dev(movie).view(T, d1 * d2), T)
    mean = mp.sum(0).cpu().numpy().reshape(d1, d2)
    np.testing.assert_allclose(mean, mean_ref, rtol=2e-6)
    if T >= 256:
        std = (npart.sum(0) / n_var).cpu().numpy().reshape(d1, d2)
        np.testing.assert_allclose(std, std_ref, rtol=2e-5)
    else:
        assert n_var == 0 and float(npart.abs().max()) == 0.0


def test_standardize_frames(ops):
    rng = np.random.default_rng(1)
    movie = rng.integers(0, 4000, size=(40, 37)).astype(np.uint16)
    mean = rng.uniform(100, 200, 37).astype(np.float32)
    std = rng.uniform(0.5, 2, 37).astype(np.float32)
    frames = np.array([5, 0, 39, 17, 5])
    out = ops.standardize_frames(dev(movie), dev(frames), dev(mean), dev(std)).cpu().numpy()
    ref = (movie[frames].astype(np.float32) - mean) / std
    np.testing.assert_array_equal(out, ref)  # same two float32 operations -> bit exact


# ------------------------------------------------------------------------------------- gram / jacobi
@pytest.mark.parametrize("n,m,batch", [(1, 10, 2), (2, 65, 3), (7, 700, 2), (25, 1000, 1), (50, 5000, 3), (60, 500, 2), (111, 300, 1)])
def test_gram_and_jacobi(ops, n, m, batch):
    rng = np.random.default_rng(n * 7 + m)
    x = (rng.standard_normal((batch, n, m)) * np.logspace(0, -3, n)[None, :, None]).astype(np.float32)
    g = ops.gram_rows(dev(x))
    ref = np.einsum("bim,bjm->bij", x.astype(np.float64), x.astype(np.float64))
    np.testing.assert_allclose(g.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())  # fp64 accumulation
    gt = ops.gram_cols(dev(np.ascontiguousarray(x.transpose(0, 2, 1))))
    np.testing.assert_allclose(gt.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    w, vecs = ops.jacobi_eigh(g.clone(), mode=0)
    w, vecs = w.cpu().numpy(), vecs.cpu().numpy().astype(np.float64)
    for b in range(batch):
        wr = np.linalg.eigvalsh(ref[b])[::-1]
        np.testing.assert_allclose(w[b], wr, rtol=1e-9, atol=1e-13 * wr[0])
        assert np.abs(vecs[b].T @ vecs[b] - np.eye(n)).max() < 5e-6  # float32 storage of the vectors
        resid = ref[b] @ vecs[b] - vecs[b] * w[b][None, :]
        assert np.abs(resid).max() < 5e-6 * wr[0]
    # float32 sweeps: eigenvalues to 1e-6 of the largest, an orthonormal basis, invariant subspaces of the top part
    w32, v32 = ops.jacobi_eigh(g.clone(), mode=0, sweeps_f32=True)
    w32, v32 = w32.cpu().numpy(), v32.cpu().numpy().astype(np.float64)
    for b in range(batch):
        wr = np.linalg.eigvalsh(ref[b])[::-1]
        np.testing.assert_allclose(w32[b], wr, rtol=0, atol=1e-5 * wr[0])
        assert np.abs(v32[b].T @ v32[b] - np.eye(n)).max() < 2e-5
        resid = ref[b] @ v32[b] - v32[b] * w32[b][None, :]
        assert np.abs(resid).max() < 2e-5 * wr[0]
    _, tm = ops.jacobi_eigh(g.clone(), mode=1)
    q = x.transpose(0, 2, 1).astype(np.float64) @ tm.cpu().numpy().astype(np.float64)
    for b in range(batch):
        assert np.abs(q[b].T @ q[b] - np.eye(n)).max() < 2e-4  # whitening of a kappa=1e3 matrix in one pass


def test_orthonormalize_rank_deficient(ops):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((2, 300, 4)).astype(np.float32)
    x = np.concatenate([a, a[:, :, :2] * 2.0, np.zeros((2, 300, 2), np.float32)], axis=2)  # rank 4 of 8 columns
    q = ops.orthonormalize_cols(dev(x)).cpu().numpy().astype(np.float64)
    for b in range(2):
        gq = q[b].T @ q[b]
        dg = np.diag(gq)
        assert np.all((np.abs(dg - 1) < 1e-4) | (dg < 1e-8))  # columns are orthonormal or (numerically) zero
        live = dg > 0.5
        assert live.sum() >= 4
        assert np.abs(gq[np.ix_(live, live)] - np.eye(live.sum())).max() < 1e-4
        proj = q[b][:, live] @ (q[b][:, live].T @ a[b].astype(np.float64))
        assert np.abs(proj - a[b]).max() < 1e-4  # span contains the original columns


def test_matmul_3xtf32(ops):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((300, 1000)) * np.exp(rng.uniform(-8, 8, (300, 1)))).astype(np.float32)
    b = (rng.standard_normal((1000, 700)) * np.exp(rng.uniform(-8, 8, (1, 700)))).astype(np.float32)
    got = ops.matmul_3xtf32(dev(a), dev(b)).cpu().numpy().astype(np.float64)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    assert np.max(np.abs(got - ref) / scale) < 2e-6  # float32-class accuracy (plain TF32 would be ~1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("m,k,n", [(300, 1000, 700), (50, 37, 101), (1650, 808, 512)])
def test_matmul_split_tf32_bf16_pairs(ops, m, k, n):
    """One TF32 + one bf16 GEMM on operands split by pmd_split_tf32_bf16: float32-class accuracy, and the split itself:
    x = hi + lo exactly, hi representable in TF32."""
    rng = np.random.default_rng(m + n)
    a = (rng.standard_normal((m, k)) * np.exp(rng.uniform(-8, 8, (m, 1)))).astype(np.float32)
    b = (rng.standard_normal((k, n)) * np.exp(rng.uniform(-8, 8, (1, n)))).astype(np.float32)
    got = ops.matmul_3xtf32_any(dev(a), dev(b)).cpu().numpy().astype(np.float64)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    assert np.max(np.abs(got - ref) / scale) < 4e-6  # float32-class accuracy (plain TF32 would be ~1e-3)
    if k % 4 == 0 and n % 4 == 0:
        x = dev(b)
        pair = ops.split_pairs(x, 0).float().cpu().numpy()
        hi = x.cpu().numpy()
        assert np.all((hi.view(np.uint32) & 0x1FFF) == 0)
        np.testing.assert_allclose(pair[:k], hi, rtol=2.0**-8)
        np.testing.assert_allclose(pair[k:], b - hi, rtol=2.0**-8, atol=1e-38)
        assert np.max(np.abs(b - hi) / np.maximum(np.abs(b), 1e-30)) <= 2.0**-11


@pytest.mark.parametrize("m,n,ld,rank", [(100, 60, 60, 60), (400, 50, 52, 50), (400, 50, 52, 17), (144, 11, 11, 11), (30, 8, 12, 5)])
def test_block_orth(ops, m, n, ld, rank):
    """Fused CholQR2: orthonormal columns spanning the input's column space; dependent columns dropped."""
    rng = np.random.default_rng(m + n)
    nbat = 5
    base = rng.standard_normal((nbat, m, rank)) * np.logspace(0, -3, rank)[None, None, :]
    mix = rng.standard_normal((nbat, rank, n))
    x = np.zeros((nbat, m, ld), np.float32)
    x[:, :, :n] = base @ mix
    x[:, :, n:] = 3.0  # padding columns must be left alone
    q = ops.block_orth(dev(x.copy()), n).cpu().numpy()
    for b in range(nbat):
        qq = q[b][:, :n].astype(np.float64)
        live = np.linalg.norm(qq, axis=0) > 0.5
        # numerically dependent columns are dropped; float32 rounding of the input can keep a few of them alive
        assert rank <= live.sum() <= n and (rank == n or live.sum() < n)
        g = qq[:, live].T @ qq[:, live]
        np.testing.assert_allclose(g, np.eye(int(live.sum())), atol=5e-6)
        assert np.all(qq[:, ~live] == 0)
        # same column space: projecting the input on span(Q) reproduces it
        xin = x[b][:, :n].astype(np.float64)
        proj = qq[:, live] @ (qq[:, live].T @ xin)
        assert np.abs(proj - xin).max() < 2e-5 * np.abs(xin).max()
        assert np.all(q[b][:, n:] == 3.0)


def test_block_orth_with_external_gram(ops):
    """X = B V^T with G = V V^T given: the result spans B * rowspace(V)^T and is orthonormal."""
    rng = np.random.default_rng(4)
    nbat, m, r, t = 3, 120, 12, 300
    bmat = rng.standard_normal((nbat, m, t))
    v = rng.standard_normal((nbat, r, t)) * np.logspace(0, -4, r)[None, :, None]
    g = v @ v.transpose(0, 2, 1)
    x = np.zeros((nbat, m, 12), np.float32)
    x[:, :, :r] = bmat @ v.transpose(0, 2, 1)
    q = ops.block_orth(dev(x.copy()), r, g_ext=dev(g)).cpu().numpy().astype(np.float64)
    for b in range(nbat):
        np.testing.assert_allclose(q[b].T @ q[b], np.eye(r), atol=5e-6)
        vb = np.linalg.qr(v[b].T)[0]  # orthonormal basis of the row space
        ref = bmat[b] @ vb
        proj = q[b] @ (q[b].T @ ref)
        assert np.abs(proj - ref).max() < 1e-3 * np.abs(ref).max()


# ------------------------------------------------------------------------------------ block kernels
def _block_setup(rng, t, d1, d2, bh, bw):
    y = rng.standard_normal((t, d1, d2)).astype(np.float32)
    starts = np.array([(k, j) for k in O.tile_starts(d1, bh) for j in O.tile_starts(d2, bw)], dtype=np.int32)
    return y, starts


def _pixel_major(y, ld=None):
    """(t, d1, d2) -> pixel-major (d, ld) float32 with zero padding columns."""
    t = y.shape[0]
    ld = (t + 3) // 4 * 4 if ld is None else ld
    yt = np.zeros((y.shape[1] * y.shape[2], ld), np.float32)
    yt[:, :t] = y.reshape(t, -1).T
    return yt


def test_standardize_frames_t(ops):
    rng = np.random.default_rng(3)
    T, d = 77, 1000
    movie = rng.integers(0, 4000, size=(T, d)).astype(np.uint16)
    mean = rng.uniform(1000, 3000, d).astype(np.float32)
    std = rng.uniform(0.5, 40, d).astype(np.float32)
    frames = rng.choice(T, 41, replace=False).astype(np.int64)
    out = ops.standardize_frames_t(dev(movie), dev(frames), dev(mean), dev(std)).cpu().numpy()
    assert out.shape == (d, 44)
    ref = ((movie[frames].astype(np.float32) - mean) / std).T
    np.testing.assert_array_equal(out[:, :41], ref)  # same float32 operations as the reference: bit exact
    assert np.all(out[:, 41:] == 0)


@pytest.mark.parametrize("bh,bw,saf,taf", [(16, 16, 2, 10), (10, 14, 2, 5), (11, 13, 2, 4), (12, 12, 3, 6), (20, 20, 2, 10)])
def test_block_pool_tavg_and_unpool(ops, bh, bw, saf, taf):
    rng = np.random.default_rng(bh * bw)
    t, d1, d2 = 120, 33, 41
    y, starts = _block_setup(rng, t, d1, d2, bh, bw)
    bta = ops.block_pool_tavg(dev(_pixel_major(y)), t, d2, dev(starts), bh, bw, saf, taf).cpu().numpy()
    pooled, bta2 = ops.block_pool_full(dev(_pixel_major(y)), t, d2, dev(starts), bh, bw, saf, taf)
    assert torch.equal(bta2.cpu(), torch.from_numpy(bta))  # same arithmetic, same order
    pooled = pooled.cpu().numpy()
    r = 3
    ph, pw = -(-bh // saf), -(-bw // saf)
    uds = rng.standard_normal((len(starts), ph * pw, r)).astype(np.float32)
    w4 = ops.block_unpool(dev(uds), bh, bw, saf, 4).cpu().numpy()
    for b, (i0, j0) in enumerate(starts):
        block = y[:, i0 : i0 + bh, j0 : j0 + bw].transpose(1, 2, 0)
        ds = O.downsample_average_pooling(block, saf)
        ta = ds.reshape(ph * pw, t // taf, taf).mean(axis=2)  # C-order pooled pixel index, consecutive frame bins
        np.testing.assert_allclose(bta[b], ta, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(pooled[b][:, :t], ds.reshape(ph * pw, t), rtol=1e-6, atol=1e-6)
        assert np.all(pooled[b][:, t:] == 0)
        lhs = w4[b][:, :r].T.astype(np.float64) @ block.reshape(bh * bw, t).astype(np.float64)
        rhs = uds[b].T.astype(np.float64) @ ds.reshape(ph * pw, t).astype(np.float64)
        np.testing.assert_allclose(lhs, rhs, atol=1e-4)
        assert np.all(w4[b][:, r:] == 0)


@pytest.mark.parametrize("bh,bw,r,t", [(16, 16, 8, 300), (20, 20, 50, 130), (10, 12, 5, 64), (40, 40, 13, 100), (22, 22, 64, 515),
                                       (32, 32, 50, 259), (10, 10, 1, 30)])
def test_block_project_and_spatial(ops, bh, bw, r, t):
    rng = np.random.default_rng(r + t)
    d1, d2 = 47, 52
    y, starts = _block_setup(rng, t, d1, d2, bh, bw)
    nb = len(starts)
    rp = (r + 3) // 4 * 4
    w = np.zeros((nb, bh * bw, rp), np.float32)
    w[:, :, :r] = rng.standard_normal((nb, bh * bw, r))
    yt = _pixel_major(y)
    ld = yt.shape[1]
    yd, sd = dev(yt), dev(starts)
    out = ops.block_project(yd, 0, ld, d2, sd, bh, bw, dev(w), r).cpu().numpy()
    assert out.shape == (nb, r, ld)
    vb = np.zeros((nb, r, ld), np.float32)
    vb[:, :, :t] = rng.standard_normal((nb, r, t))
    s = ops.block_spatial(yd, 0, ld, d2, sd, bh, bw, dev(vb), rp).cpu().numpy()
    for b, (i0, j0) in enumerate(starts):
        blk = y[:, i0 : i0 + bh, j0 : j0 + bw].reshape(t, bh * bw).T.astype(np.float64)  # (b, t)
        ref = w[b][:, :r].T.astype(np.float64) @ blk
        np.testing.assert_allclose(out[b][:, :t], ref, rtol=1e-4, atol=2e-5 * np.abs(ref).max())
        assert np.all(out[b][:, t:] == 0)
        ref_s = blk @ vb[b][:, :t].T.astype(np.float64)
        np.testing.assert_allclose(s[b][:, :r], ref_s, rtol=1e-4, atol=2e-5 * np.abs(ref_s).max())
        assert np.all(s[b][:, r:] == 0)


@pytest.mark.parametrize("bh,bw,r,t", [(20, 20, 50, 300), (16, 16, 8, 130), (10, 12, 5, 64), (22, 22, 64, 515), (10, 10, 1, 30)])
def test_block_project_tensor_core(ops, bh, bw, r, t):
    """tcgen05 3xTF32 block projection against float64 (float32-class accuracy, not TF32 accuracy)."""
    rng = np.random.default_rng(r + t)
    d1, d2 = 47, 52
    y, starts = _block_setup(rng, t, d1, d2, bh, bw)
    y *= np.exp(rng.uniform(-3, 3, size=(1, d1, d2))).astype(np.float32)  # wide dynamic range across pixels
    nb = len(starts)
    rp = (r + 3) // 4 * 4
    w = np.zeros((nb, bh * bw, rp), np.float32)
    w[:, :, :r] = rng.standard_normal((nb, bh * bw, r))
    yt = _pixel_major(y)
    ld = yt.shape[1]
    out = ops.block_project_tc(dev(yt), 0, ld, d2, dev(starts), bh, bw, dev(w), r).cpu().numpy()
    assert out.shape == (nb, r, ld)
    for b, (i0, j0) in enumerate(starts):
        blk = y[:, i0 : i0 + bh, j0 : j0 + bw].reshape(t, bh * bw).T.astype(np.float64)
        ref = w[b][:, :r].T.astype(np.float64) @ blk
        scale = np.abs(w[b][:, :r].T.astype(np.float64)) @ np.abs(blk)
        assert np.max(np.abs(out[b][:, :t] - ref) / scale) < 3e-6
        assert np.all(out[b][:, t:] == 0)


@pytest.mark.parametrize("bh,bw,r,t", [(20, 20, 50, 300), (16, 16, 8, 130), (10, 12, 5, 64), (22, 22, 64, 515), (32, 32, 50, 259)])
def test_block_spatial_tensor_core(ops, bh, bw, r, t):
    """tcgen05 3xTF32 spatial projection against float64."""
    rng = np.random.default_rng(r + t)
    d1, d2 = 47, 52
    y, starts = _block_setup(rng, t, d1, d2, bh, bw)
    y *= np.exp(rng.uniform(-3, 3, size=(1, d1, d2))).astype(np.float32)
    nb = len(starts)
    rp = (r + 3) // 4 * 4
    yt = _pixel_major(y)
    ld = yt.shape[1]
    vb = np.zeros((nb, r, ld), np.float32)
    vb[:, :, :t] = rng.standard_normal((nb, r, t)) * np.exp(rng.uniform(-2, 2, size=(nb, r, 1)))
    s = ops.block_spatial_tc(dev(yt), 0, ld, d2, dev(starts), bh, bw, dev(vb), rp).cpu().numpy()
    for b, (i0, j0) in enumerate(starts):
        blk = y[:, i0 : i0 + bh, j0 : j0 + bw].reshape(t, bh * bw).T.astype(np.float64)  # (b, t)
        ref = blk @ vb[b][:, :t].T.astype(np.float64)
        scale = np.abs(blk) @ np.abs(vb[b][:, :t].T.astype(np.float64))
        assert np.max(np.abs(s[b][:, :r] - ref) / scale) < 3e-6
        assert np.all(s[b][:, r:] == 0)


def test_block_project_batched_movies(ops):
    """movie_batch_stride != 0: every 'block' is its own small pixel-major movie (threshold simulation)."""
    rng = np.random.default_rng(0)
    m, t, bh, bw, r = 5, 90, 12, 10, 3
    ld = 92
    movies = np.zeros((m, bh * bw, ld), np.float32)
    movies[:, :, :t] = rng.standard_normal((m, bh * bw, t))
    w = np.zeros((m, bh * bw, 4), np.float32)
    w[:, :, :r] = rng.standard_normal((m, bh * bw, r))
    starts = np.zeros((m, 2), np.int32)
    out = ops.block_project(dev(movies), bh * bw * ld, ld, bw, dev(starts), bh, bw, dev(w), r).cpu().numpy()
    ref = np.einsum("mqc,mqt->mct", w[:, :, :r].astype(np.float64), movies.astype(np.float64))
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("mcf", [1, 2, 3])
def test_block_stats_rank(ops, mcf):
    rng = np.random.default_rng(mcf)
    nb, bh, bw, r, t = 7, 12, 14, 9, 257
    rp = 12
    ii, jj = np.mgrid[0:bh, 0:bw]
    u = np.zeros((nb, bh * bw, rp), np.float32)
    v = np.zeros((nb, r, t), np.float32)
    for b in range(nb):
        for c in range(r):
            smooth = rng.uniform() < 0.6
            img = np.exp(-((ii - rng.uniform(0, bh)) ** 2 + (jj - rng.uniform(0, bw)) ** 2) / 20.0) if smooth else rng.standard_normal((bh, bw))
            tr = np.cumsum(rng.standard_normal(t)) if smooth else rng.standard_normal(t)
            u[b, :, c] = img.reshape(-1)
            v[b, c] = tr
    v[3, 4] = 0.0  # NaN statistic -> failure
    thr_s, thr_t = 0.9, 1.5
    vpad = np.concatenate([v, rng.standard_normal((nb, r, 3)).astype(np.float32)], axis=2)  # padding must be ignored
    ss, ts, ranks = ops.block_stats_rank(dev(u), dev(vpad), bh, bw, r, thr_s, thr_t, mcf, t=t)
    ss, ts, ranks = ss.cpu().numpy(), ts.cpu().numpy(), ranks.cpu().numpy()
    for b in range(nb):
        u3 = u[b][:, :r].reshape(bh, bw, r)
        good, ss_ref, ts_ref = O.fitness_decisions(u3, v[b], thr_s, thr_t)
        np.testing.assert_allclose(ss[b], ss_ref, rtol=1e-5)
        np.testing.assert_allclose(ts[b], ts_ref, rtol=1e-5, equal_nan=True)
        assert ranks[b] == O.filter_by_failures(good > 0, mcf).sum()


def test_assemble_u(ops):
    rng = np.random.default_rng(2)
    d1, d2, bh, bw, rp = 30, 26, 12, 10, 8
    starts = np.array([(k, j) for k in O.tile_starts(d1, bh) for j in O.tile_starts(d2, bw)], dtype=np.int32)
    nb = len(starts)
    u = rng.standard_normal((nb, bh * bw, rp)).astype(np.float32)
    ranks = rng.integers(1, 7, nb).astype(np.int32)
    col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]]).astype(np.int64)
    wts = O.pyramid_weights(bh, bw)
    cumw = np.zeros((d1, d2))
    for k, j in starts:
        cumw[k : k + bh, j : j + bw] += wts
    uv64, uv32 = ops.assemble_u(dev(u), bh, bw, dev(starts), dev(ranks), dev(col0), dev(wts.reshape(-1)), dev(cumw.reshape(-1)),
                                d2, int(ranks.sum()))
    uv64 = uv64.cpu().numpy()
    for b, (i0, j0) in enumerate(starts):
        for c in range(ranks[b]):
            ref = (1.0 / cumw[i0 : i0 + bh, j0 : j0 + bw]) * (u[b, :, c].reshape(bh, bw).astype(np.float64) * wts.astype(np.float64))
            np.testing.assert_array_equal(uv64[col0[b] + c].reshape(bh, bw), ref)  # same float64 operations
    np.testing.assert_array_equal(uv32.cpu().numpy(), uv64.astype(np.float32))


# ------------------------------------------------------------------------------------------- K7 / K9
def _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K):
    starts = np.array([(k, j) for k in O.tile_starts(d1, bh) for j in O.tile_starts(d2, bw)], dtype=np.int32)
    nb = len(starts)
    ranks = rng.integers(1, max_rank + 1, nb).astype(np.int32)
    col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]]).astype(np.int64)
    n_local = int(ranks.sum())
    uv = rng.standard_normal((n_local, bh * bw)).astype(np.float32)
    bg = rng.standard_normal((K, d1 * d2)).astype(np.float32)
    rows, cols, vals = [], [], []
    qi, qj = np.divmod(np.arange(bh * bw), bw)
    for b, (i0, j0) in enumerate(starts):
        pix = (i0 + qi) * d2 + j0 + qj
        for c in range(ranks[b]):
            rows.append(pix), cols.append(np.full(bh * bw, col0[b] + c)), vals.append(uv[col0[b] + c])
    for k in range(K):
        rows.append(np.arange(d1 * d2)), cols.append(np.full(d1 * d2, n_local + k)), vals.append(bg[k])
    U = sp.csr_matrix((np.concatenate(vals).astype(np.float64), (np.concatenate(rows), np.concatenate(cols))), shape=(d1 * d2, n_local + K))
    return starts, ranks, col0, uv, bg, U


@pytest.mark.parametrize(
    "bh,bw,max_rank,dtype",
    [(10, 10, 3, np.float32), (16, 16, 9, np.uint16), (20, 20, 5, np.float32), (22, 22, 2, np.int16), (32, 32, 6, np.float32),
     (40, 40, 3, np.uint8)],
)
def test_project_local_and_dense(ops, bh, bw, max_rank, dtype):
    rng = np.random.default_rng(bh + max_rank)
    d1, d2, T, K = 61, 83, 777, 5
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
    y = rng.uniform(0, 200, size=(T, d1 * d2))
    movie = (np.rint(y) if np.issubdtype(dtype, np.integer) else y).astype(dtype)
    mean = rng.uniform(80, 120, d1 * d2).astype(np.float32)
    std = rng.uniform(0.5, 2, d1 * d2).astype(np.float32)
    inv = (1.0 / std).astype(np.float32)
    n_local = int(ranks.sum())
    z = torch.full((n_local + K, T + 3), 7.0, dtype=torch.float32, device="cuda")
    if bh * bw > 512:
        z[:n_local].zero_()
    z[n_local:].zero_()
    tasks = dev(ops.make_tasks(ranks))
    ops.project_local(dev(movie), d2, dev(starts), bh, bw, dev(ranks), dev(col0), tasks, dev(uv), dev(mean), dev(inv), z[:n_local])
    ops.project_dense(dev(movie), dev(bg), dev(mean), dev(inv), z[n_local:])
    yc = (movie.astype(np.float64) - mean) / std
    ref = (U.T @ yc.T)  # (R, T)
    got = z.cpu().numpy()
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got[:, :T], ref, rtol=0, atol=2e-5 * scale)  # float32 dot products of length <= d
    assert np.all(got[:, T:] == np.where(np.arange(n_local + K)[:, None] < n_local, 0.0 if bh * bw > 512 else 7.0, 0.0))


@pytest.mark.parametrize(
    "bh,bw,max_rank,dtype,d1,d2",
    [(10, 10, 3, np.float32, 61, 83), (16, 16, 9, np.uint16, 61, 83), (20, 20, 5, np.float32, 112, 95), (22, 22, 13, np.int16, 61, 83),
     (20, 12, 6, np.float32, 64, 40), (20, 20, 2, np.float64, 20, 20)],
)
def test_project_supertile(ops, bh, bw, max_rank, dtype, d1, d2):
    rng = np.random.default_rng(bh * 3 + max_rank)
    T, K = 1100, 1
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
    y = rng.uniform(0, 200, size=(T, d1 * d2))
    movie = (np.rint(y) if np.issubdtype(dtype, np.integer) else y).astype(dtype)
    mean = rng.uniform(80, 120, d1 * d2).astype(np.float32)
    std = rng.uniform(0.5, 2, d1 * d2).astype(np.float32)
    inv = (1.0 / std).astype(np.float32)
    n_local = int(ranks.sum())
    st = ops.make_supertiles(O.tile_starts(d1, bh), O.tile_starts(d2, bw), bh, bw, ranks, col0)
    std_ = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    z = torch.full((n_local, T + 5), 7.0, dtype=torch.float32, device="cuda")
    ops.project_supertile(dev(movie), d2, std_, bh, bw, dev(uv), dev(mean), dev(inv), z)
    yc = (movie.astype(np.float32).astype(np.float64) - mean) / std
    ref = (U.T @ yc.T)[:n_local]
    got = z.cpu().numpy()
    np.testing.assert_allclose(got[:, :T], ref, rtol=0, atol=2e-5 * np.abs(ref).max())
    assert np.all(got[:, T:] == 7.0)  # nothing written beyond the movie
    z2 = torch.zeros((n_local, T), dtype=torch.float32, device="cuda")
    if dtype == np.float32:
        ops.project_supertile(dev(movie), d2, std_, bh, bw, dev(uv), None, None, z2)
        ref2 = (U.T @ movie.astype(np.float64).T)[:n_local]
        np.testing.assert_allclose(z2.cpu().numpy(), ref2, rtol=0, atol=2e-5 * np.abs(ref2).max())


@pytest.mark.parametrize("bh,bw,d1,d2,max_rank,mcols", [(20, 20, 112, 95, 5, 9), (10, 14, 33, 47, 3, 9), (16, 16, 16, 16, 4, 9),
                                                          (12, 12, 50, 37, 7, 9), (12, 12, 50, 37, 30, 70), (20, 20, 112, 95, 5, 300)])
def test_utu_gram_times(ops, bh, bw, d1, d2, max_rank, mcols):
    """Block-sparse U^T U (pmd_utu_pairs + pmd_utu_apply_tiles + pmd_project_cols_f64) applied to a dense right factor:
    ranks above 24 (several row chunks per block), more than 256 columns (several column CTAs), ragged edges."""
    from localmd_b200.decomposition import SparseU

    rng = np.random.default_rng(bh + d1)
    K = 3
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
    uv64 = uv.astype(np.float64)
    su = SparseU(starts, dev(starts), bh, bw, d1, d2, ranks.astype(np.int64), dev(ranks), dev(uv64), dev(uv), dev(bg))
    right = rng.standard_normal((U.shape[1], mcols))
    got = su.utu_times_f64(dev(right)).cpu().numpy()
    ref = (U.T @ U) @ right
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())
    rowptr, cols, vals = su.gram()[0]
    n_local = int(ranks.sum())
    L = sp.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), rowptr.cpu().numpy()), shape=(n_local, n_local))
    assert L.has_canonical_format
    np.testing.assert_allclose(L.toarray(), (U.T @ U).toarray()[:n_local, :n_local], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize(
    "bh,bw,max_rank,dtype,d1,d2,K,T",
    [(10, 10, 3, np.float32, 61, 83, 1, 300), (16, 16, 9, np.uint16, 61, 83, 5, 1100), (20, 20, 13, np.float32, 112, 95, 15, 777),
     (22, 22, 20, np.int16, 61, 83, 9, 258), (20, 12, 6, np.float32, 64, 40, 16, 513), (20, 20, 2, np.float64, 20, 20, 2, 64),
     (32, 32, 6, np.uint8, 70, 96, 3, 260), (40, 40, 11, np.float32, 90, 101, 4, 255)],
)
def test_project_stream(ops, bh, bw, max_rank, dtype, d1, d2, K, T):
    """K7 v3 (strip-streaming kernel): local + dense columns in one pass, against float64 U^T Y."""
    rng = np.random.default_rng(bh * 3 + max_rank)
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
    y = rng.uniform(0, 200, size=(T, d1 * d2))
    movie = (np.rint(y) if np.issubdtype(dtype, np.integer) else y).astype(dtype)
    mean = rng.uniform(80, 120, d1 * d2).astype(np.float32)
    std = rng.uniform(0.5, 2, d1 * d2).astype(np.float32)
    inv = (1.0 / std).astype(np.float32)
    n_local = int(ranks.sum())
    st = ops.make_strips(O.tile_starts(d1, bh), O.tile_starts(d2, bw), bh, bw, d1, d2, ranks, col0, K)
    assert st is not None
    std_ = {k: (dev(v) if k in ("items", "slot_ptr", "tasks") else v) for k, v in st.items()}
    upack = ops.pack_strip_u(st, dev(uv), dev(bg), bh * bw)
    z = torch.full((n_local + K, T + 5), 7.0, dtype=torch.float32, device="cuda")
    ops.project_stream(dev(movie), d2, std_, upack, dev(mean), dev(inv), z[:n_local], z[n_local:])
    yc = (movie.astype(np.float32).astype(np.float64) - mean) / std
    ref = U.T @ yc.T
    got = z.cpu().numpy()
    np.testing.assert_allclose(got[:, :T], ref, rtol=0, atol=2e-5 * np.abs(ref).max())
    assert np.all(got[:, T:] == 7.0)  # nothing written beyond the movie
    if dtype == np.float32:
        z2 = torch.zeros((n_local + K, T), dtype=torch.float32, device="cuda")
        ops.project_stream(dev(movie), d2, std_, upack, None, None, z2[:n_local], z2[n_local:])
        ref2 = U.T @ movie.astype(np.float64).T
        np.testing.assert_allclose(z2.cpu().numpy(), ref2, rtol=0, atol=2e-5 * np.abs(ref2).max())


@pytest.mark.parametrize(
    "bh,bw,max_rank,dtype,d1,d2,K,T,G",
    [(20, 20, 13, np.float32, 112, 96, 15, 700, None), (10, 10, 3, np.uint16, 61, 84, 1, 300, None),
     (16, 16, 12, np.float32, 70, 96, 5, 1100, 2), (22, 22, 20, np.int16, 61, 88, 9, 258, None),
     (20, 12, 6, np.float32, 64, 40, 16, 513, None), (20, 20, 2, np.float64, 24, 32, 2, 64, None),
     (32, 32, 6, np.uint8, 70, 96, 3, 260, None), (40, 40, 11, np.float32, 90, 104, 4, 255, None),
     (20, 20, 50, np.float32, 60, 80, 0, 130, None), (20, 20, 4, np.int32, 50, 60, 2, 129, 1)],
)
def test_project_stream_tc(ops, bh, bw, max_rank, dtype, d1, d2, K, T, G):
    """K7 on tcgen05 (TF32 main product + bf16 correction MMA, tensor-memory slot accumulators): local + dense
    columns in one pass, against float64 U^T Y.  Tolerance as for the SIMT kernel (float32-accurate)."""
    rng = np.random.default_rng(bh * 5 + max_rank)
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
    y = rng.uniform(0, 200, size=(T, d1 * d2))
    movie = (np.rint(y) if np.issubdtype(dtype, np.integer) else y).astype(dtype)
    mean = rng.uniform(80, 120, d1 * d2).astype(np.float32)
    std = rng.uniform(0.5, 2, d1 * d2).astype(np.float32)
    inv = (1.0 / std).astype(np.float32)
    n_local = int(ranks.sum())
    st = ops.make_strips_tc(O.tile_starts(d1, bh), O.tile_starts(d2, bw), bh, bw, d1, d2, ranks, col0, K, G=G)
    assert st is not None
    std_ = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    bgd = dev(bg) if K else None
    bimg = ops.pack_strips_tc(std_, dev(uv), bgd, bh * bw, d2)
    z = torch.full((n_local + K, T + 5), 7.0, dtype=torch.float32, device="cuda")
    mv = dev(movie)
    assert ops.project_stream_tc_ok(mv, d2, dev(mean), dev(inv))
    ops.project_stream_tc(mv, d2, std_, bimg, dev(mean), dev(inv), z[:n_local], z[n_local:])
    yc = (movie.astype(np.float32).astype(np.float64) - mean) / std
    ref = U.T @ yc.T
    got = z.cpu().numpy()
    np.testing.assert_allclose(got[:, :T], ref[: n_local + K], rtol=0, atol=2e-5 * np.abs(ref).max())
    assert np.all(got[:, T:] == 7.0)  # nothing written beyond the movie
    if dtype == np.float32:
        z2 = torch.zeros((n_local + K, T), dtype=torch.float32, device="cuda")
        ops.project_stream_tc(mv, d2, std_, bimg, None, None, z2[:n_local], z2[n_local:])
        ref2 = U.T @ movie.astype(np.float64).T
        np.testing.assert_allclose(z2.cpu().numpy(), ref2[: n_local + K], rtol=0, atol=2e-5 * np.abs(ref2).max())


@pytest.mark.parametrize(
    "bh,bw,max_rank,dtype,d1,d2,K,T,W,N",
    [(20, 20, 9, np.float32, 60, 128, 15, 700, None, None), (20, 20, 9, np.uint16, 50, 104, 15, 300, 32, 96), (20, 20, 9, np.uint16, 50, 100, 15, 300, 32, 96),
     (16, 16, 12, np.float32, 70, 96, 5, 1100, 64, 128), (22, 22, 20, np.int16, 61, 88, 9, 258, None, None),
     (20, 12, 6, np.float32, 64, 40, 16, 513, None, 192), (20, 20, 2, np.float64, 24, 32, 2, 64, None, None),
     (32, 32, 6, np.uint8, 70, 96, 3, 260, 64, None), (40, 40, 11, np.float32, 90, 104, 4, 255, None, None),
     (20, 20, 50, np.float32, 60, 80, 0, 130, None, None), (20, 20, 4, np.int32, 50, 60, 2, 129, 128, 96),
     (40, 40, 30, np.float32, 128, 256, 15, 400, None, None), (20, 20, 6, np.float32, 128, 500, 15, 385, None, None)],
)
def test_project_stream_ts(ops, bh, bw, max_rank, dtype, d1, d2, K, T, W, N):
    """K7 with TMA-fed raw tiles and the movie operand in tensor memory (tcgen05 TS form), exact-partition strips with
    atomically added partial sums of blocks shared by two strips, 1/std folded into the coefficient images: local + dense
    columns in one pass, against float64 U^T Y.  Tolerance as for the other K7 kernels (float32-accurate)."""
    rng = np.random.default_rng(bh * 5 + max_rank)
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
    y = rng.uniform(0, 200, size=(T, d1 * d2))
    movie = (np.rint(y) if np.issubdtype(dtype, np.integer) else y).astype(dtype)
    mean = rng.uniform(80, 120, d1 * d2).astype(np.float32)
    std = rng.uniform(0.5, 2, d1 * d2).astype(np.float32)
    inv = (1.0 / std).astype(np.float32)
    n_local = int(ranks.sum())
    st = ops.make_strips_ts(O.tile_starts(d1, bh), O.tile_starts(d2, bw), bh, bw, d1, d2, ranks, col0, K, W=W, N=N)
    assert st is not None
    if W:
        assert st["W"] == W
    if N:
        assert st["N"] == N
    std_ = {k: (dev(v) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
    bgd = dev(bg) if K else None
    bimg = ops.pack_strips_ts(std_, dev(uv), bgd, dev(inv), bh * bw, d2)
    z = torch.full((n_local + K, T + 5), 7.0, dtype=torch.float32, device="cuda")
    mv = dev(movie)
    if not ops.project_stream_ts_ok(mv, d2, dev(mean)):
        with pytest.raises(Exception):   # rejected by the entry point as well (callers fall back to the older kernels)
            ops.project_stream_ts(mv, d2, std_, bimg, dev(mean), z[:n_local], z[n_local:])
        return
    ops.project_stream_ts(mv, d2, std_, bimg, dev(mean), z[:n_local], z[n_local:])
    yc = (movie.astype(np.float32).astype(np.float64) - mean) / std
    ref = U.T @ yc.T
    got = z.cpu().numpy()
    np.testing.assert_allclose(got[:, :T], ref[: n_local + K], rtol=0, atol=2e-5 * np.abs(ref).max())
    assert np.all(got[:, T:] == 7.0)  # nothing written beyond the movie
    if dtype == np.float32:
        bimg2 = ops.pack_strips_ts(std_, dev(uv), bgd, None, bh * bw, d2)
        z2 = torch.zeros((n_local + K, T), dtype=torch.float32, device="cuda")
        ops.project_stream_ts(mv, d2, std_, bimg2, None, z2[:n_local], z2[n_local:])
        ref2 = U.T @ movie.astype(np.float64).T
        np.testing.assert_allclose(z2.cpu().numpy(), ref2[: n_local + K], rtol=0, atol=2e-5 * np.abs(ref2).max())


def test_project_without_standardisation(ops):
    rng = np.random.default_rng(9)
    d1, d2, T, K, bh, bw = 40, 36, 50, 2, 16, 16
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, 4, K)
    movie = rng.standard_normal((T, d1 * d2)).astype(np.float32)
    n_local = int(ranks.sum())
    z = torch.zeros((n_local + K, T), dtype=torch.float32, device="cuda")
    ops.project_local(dev(movie), d2, dev(starts), bh, bw, dev(ranks), dev(col0), dev(ops.make_tasks(ranks)), dev(uv), None, None, z[:n_local])
    ops.project_dense(dev(movie), dev(bg), None, None, z[n_local:])
    ref = U.T @ movie.astype(np.float64).T
    np.testing.assert_allclose(z.cpu().numpy(), ref, atol=2e-5 * np.abs(ref).max())


@pytest.mark.parametrize("n", [1, 3, 4, 10])
def test_reconstruct(ops, n):
    rng = np.random.default_rng(n)
    d1, d2 = 23, 31
    starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, 10, 12, 3, 2)
    R = U.shape[1]
    c = rng.standard_normal((R, n)).astype(np.float32)
    pix = rng.permutation(d1 * d2)[:200].astype(np.int32)
    scale = rng.uniform(0.5, 2, d1 * d2).astype(np.float32)
    shift = rng.uniform(100, 200, d1 * d2).astype(np.float32)
    U32 = U.astype(np.float32)
    out = ops.reconstruct(dev(U32.indptr.astype(np.int64)), dev(U32.indices.astype(np.int32)), dev(U32.data), dev(c), dev(pix),
                          dev(scale), dev(shift)).cpu().numpy()
    ref = (U[pix] @ c.astype(np.float64)).T * scale[pix][None] + shift[pix][None]
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-4)
    out2 = ops.reconstruct(dev(U32.indptr.astype(np.int64)), dev(U32.indices.astype(np.int32)), dev(U32.data), dev(c), dev(pix),
                           None, None).cpu().numpy()
    np.testing.assert_allclose(out2, (U[pix] @ c.astype(np.float64)).T, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("d,t,K", [(300, 1000, 15), (1030, 37 * 4, 1), (257, 1028, 16), (64, 8, 3)])
def test_bg_filter_t(ops, d, t, K):
    """Background projection / removal on the pixel-major init movie (csrc/bgfilter.cu) against float64 matmuls."""
    rng = np.random.default_rng(d + K)
    ld = (t + 3) // 4 * 4
    yt = np.zeros((d, ld), np.float32)
    yt[:, :t] = rng.standard_normal((d, t)).astype(np.float32)
    bg = np.linalg.qr(rng.standard_normal((d, K)))[0].T.astype(np.float32)
    ytd, bgd = dev(yt), dev(bg)
    vbg = ops.bg_project_t(ytd, bgd, n_ranges=7)
    ref_v = bg.astype(np.float64) @ yt.astype(np.float64)
    np.testing.assert_allclose(vbg.cpu().numpy(), ref_v, rtol=0, atol=2e-5 * np.abs(ref_v).max())
    ops.bg_remove_t(ytd, bgd, vbg)
    ref_y = yt.astype(np.float64) - bg.T.astype(np.float64) @ vbg.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(ytd.cpu().numpy(), ref_y, rtol=0, atol=1e-5)
    assert np.all(ytd.cpu().numpy()[:, t:] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("d,n,l", [(700, 1000, 25), (333, 130, 11), (64, 1500, 32), (5000, 37, 16)])
def test_rows_sketch(ops, d, n, l):
    """y = yt[:, :n] @ omega on the pixel-major frames (csrc/bgbasis.cu) against a float64 matmul; padding columns of yt
    hold garbage and must not leak into the result."""
    rng = np.random.default_rng(d + n)
    ld = (n + 3) // 4 * 4 + 4
    yt = rng.standard_normal((d, ld)).astype(np.float32)
    yt[:, n:] = np.nan
    om = rng.standard_normal((n, l)).astype(np.float32)
    y = ops.rows_sketch(dev(yt), n, dev(om)).cpu().numpy()
    ref = yt[:, :n].astype(np.float64) @ om.astype(np.float64)
    np.testing.assert_allclose(y, ref, rtol=0, atol=3e-6 * np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("batch,d,ldx,k,nc,transposed", [(1, 1000, 25, 25, 25, False), (1, 777, 28, 25, 15, True), (3, 130, 32, 32, 32, False),
                                                          (2, 65, 7, 5, 3, True)])
def test_rows_times_small(ops, batch, d, ldx, k, nc, transposed):
    rng = np.random.default_rng(d + k)
    x = rng.standard_normal((batch, d, ldx)).astype(np.float32)
    m = rng.standard_normal((batch, k, nc)).astype(np.float32)
    out = ops.rows_times_small(dev(x), dev(m), transposed=transposed).cpu().numpy()
    ref = np.einsum("bpj,bjc->bpc", x[:, :, :k].astype(np.float64), m.astype(np.float64))
    if transposed:
        ref = ref.transpose(0, 2, 1)
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6 * np.abs(ref).max())


@pytest.mark.gpu
def test_bg_project_t_wide(ops):
    """The coefficient pass of the background rSVD: pmd_bg_project_t with 17 <= k <= 32 rows."""
    rng = np.random.default_rng(3)
    d, t, K = 3000, 1000, 25
    yt = rng.standard_normal((d, t)).astype(np.float32)
    q = np.linalg.qr(rng.standard_normal((d, K)))[0].T.astype(np.float32)
    v = ops.bg_project_t(dev(yt), dev(q), n_ranges=11).cpu().numpy()
    ref = q.astype(np.float64) @ yt.astype(np.float64)
    np.testing.assert_allclose(v, ref, rtol=0, atol=2e-5 * np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("K", [3, 15])
def test_background_basis_matches_oracle(ops, K):
    """background_basis (standardise + transpose, streaming sketch, CholQR-style orthonormalisation, coefficient pass,
    rotation) spans the same subspace as the oracle's randomised SVD with the same sketch matrix (pmd_loader.py:46-68)."""
    import oracle.pmd_oracle as O
    from localmd_b200 import decomposition as D
    from localmd_b200.dataset import DeviceMovie

    from synth import make_movie

    T, d1, d2 = 600, 24, 28
    movie = make_movie(T, d1, d2, n_cells=6, seed=2)
    rng = np.random.default_rng(K)
    frames = sorted(rng.choice(T, 200, replace=False).tolist())
    sketch = rng.standard_normal((len(frames), K + 10)).astype(np.float32)
    dm = DeviceMovie(movie, torch.device("cuda"))
    mean, std = D.compute_mean_and_noise(dm)
    bg = D.background_basis(dm, mean, std, frames, dev(sketch), K).cpu().numpy()     # (K, d)
    assert bg.shape == (K, d1 * d2)
    np.testing.assert_allclose(bg @ bg.T, np.eye(K), atol=2e-5)
    with O.precision(np.float64):
        u = O.background_basis(movie, mean.cpu().numpy().reshape(d1, d2), std.cpu().numpy().reshape(d1, d2), frames, sketch, K,
                               order="C").astype(np.float64)                                # (d, K), physical pixel order
    sv = np.linalg.svd(bg.astype(np.float64) @ u, compute_uv=False)
    assert np.arccos(np.clip(sv.min(), -1, 1)) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("n,batch", [(25, 1), (32, 3), (5, 2)])
def test_chol_whiten(ops, n, batch):
    """x @ L^-T is orthonormal for g = x^T x = L L^T; a dependent column gives a zero column."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((batch, 400, n)) * np.logspace(0, -3, n)[None, None, :]
    x[-1, :, n - 2] = x[-1, :, 0] * 2.0 - x[-1, :, 1]          # dependent column in the last matrix
    g = np.einsum("bmi,bmj->bij", x, x)
    t = ops.chol_whiten(dev(g, torch.float64)).cpu().numpy().astype(np.float64)
    q = np.einsum("bmi,bij->bmj", x, t)
    qq = np.einsum("bmi,bmj->bij", q, q)
    want = np.stack([np.eye(n)] * batch)
    want[-1, n - 2, n - 2] = 0.0
    np.testing.assert_allclose(qq, want, atol=5e-5)
    assert np.all(t[-1][:, n - 2] == 0)
    assert np.allclose(np.tril(t[0], -1), 0)



# ------------------------------------------------------------------------------------------------ symmetric float64 products
@pytest.mark.parametrize("n,k,layout,dt", [(37, 50, 0, np.float32), (130, 1000, 0, np.float32), (300, 2500, 0, np.float64),
                                           (257, 777, 1, np.float32), (129, 4099, 1, np.float64), (1, 3, 0, np.float32)])
def test_sym_product_f64(ops, n, k, layout, dt):
    """pmd_sym_product_f64 against NumPy float64: Gram a a^T (layout 0) and a^T (S a) with S symmetric (layout 1)."""
    rng = np.random.default_rng(n + k)
    if layout == 0:
        a = rng.standard_normal((n, k)).astype(dt)
        got = ops.sym_product_f64(dev(a)).cpu().numpy()
        want = a.astype(np.float64) @ a.astype(np.float64).T
    else:
        a = rng.standard_normal((k, n)).astype(dt)
        s = rng.standard_normal((k, k))
        b = (s + s.T) @ a.astype(np.float64)
        got = ops.sym_product_f64(dev(a), dev(b), layout=1).cpu().numpy()
        want = a.astype(np.float64).T @ b
    assert np.array_equal(got, got.T)                                   # both triangles from the same tiles
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * np.abs(want).max())   # float64 accumulation, reordered sums
    if layout == 0:
        again = ops.sym_product_f64(dev(a)).cpu().numpy()
        assert np.array_equal(got, again)                               # deterministic


# ------------------------------------------------------------------------------------------------ CSR export
@pytest.mark.parametrize("d1,d2,bh,bw,K,order", [(50, 44, 16, 12, 3, "F"), (40, 36, 16, 16, 2, "C"), (23, 31, 10, 14, 1, "F")])
def test_export_csr_matches_sorted_coordinate_form(ops, d1, d2, bh, bw, K, order):
    """pmd_export_csr (direct CSR from the block-component form) against the sorted coordinate form: identical indptr /
    indices / values in both row numberings, exact zeros dropped, shifted last tiles (up to 3 x 3 covering blocks)."""
    from localmd_b200.decomposition import SparseU, tile_starts

    rng = np.random.default_rng(d1 * d2)
    rows, cols = tile_starts(d1, bh), tile_starts(d2, bw)
    starts = np.array([(r, c) for r in rows for c in cols], dtype=np.int32)
    nb = len(starts)
    ranks = rng.integers(0, 5, nb).astype(np.int64)
    n_local = int(ranks.sum())
    uv = rng.standard_normal((n_local, bh * bw))
    uv[rng.random(uv.shape) < 0.1] = 0.0                      # exact zeros must be dropped
    bg = rng.standard_normal((K, d1 * d2)).astype(np.float32)
    bg[rng.random(bg.shape) < 0.05] = 0.0
    uv64 = dev(uv)
    su = SparseU(starts, dev(starts), bh, bw, d1, d2, ranks, dev(ranks.astype(np.int32)), uv64, uv64.to(torch.float32), dev(bg))
    row_ids = dev(np.arange(d1 * d2).reshape((d1, d2), order=order).reshape(-1))
    (ip, ix, v), (ip32, ix32, v32) = SparseU.finish_export(su.export_csr(row_ids))
    wp, wx, wv = su.csr(row_ids)
    pp, px, pv = su.csr_physical32()
    assert torch.equal(ip, wp) and torch.equal(ix, wx) and torch.equal(v, wv)
    assert torch.equal(ip32, pp) and torch.equal(ix32, px) and torch.equal(v32, pv)
