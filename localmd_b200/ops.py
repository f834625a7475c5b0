"""Thin torch-tensor wrappers around the C ABI (include/pmd_sm100.h).

PyTorch is plumbing only here: it owns device memory and streams; every wrapper validates shapes /
dtypes / contiguity, allocates outputs and enqueues one kernel (or a short fixed sequence) on the
current CUDA stream through ctypes.  No wrapper has a CPU path."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._tables import welch_tables

PMD_DTYPES = {
    torch.float32: 0,
    torch.uint16: 1,
    torch.int16: 2,
    torch.uint8: 3,
    torch.float64: 4,
    torch.int32: 5,
}
NUMPY_NATIVE = {np.dtype(k): v for k, v in [("float32", torch.float32), ("uint16", torch.uint16), ("int16", torch.int16),
                                               ("uint8", torch.uint8), ("float64", torch.float64), ("int32", torch.int32)]}

LAUNCHES = {"count": 0, "by_name": {}}  # number of libpmd kernel launches issued (bench.py reports it)
_KERNELS_PER_CALL = {"pmd_block_stats_rank": 3}


def _count(name, n=None):
    k = _KERNELS_PER_CALL.get(name, 1) if n is None else n
    LAUNCHES["count"] += k
    LAUNCHES["by_name"][name] = LAUNCHES["by_name"].get(name, 0) + k


def _p(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (there is no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _call(name, *args):
    rc = getattr(_lib.lib(), name)(*args)
    _lib.check(rc, name)
    _count(name)


_table_cache = {}


def _tables(device):
    key = str(device)
    if key not in _table_cache:
        tc, ts = welch_tables()
        _table_cache[key] = (torch.from_numpy(tc).to(device), torch.from_numpy(ts).to(device))
    return _table_cache[key]


def movie_dtype_code(t):
    if t.dtype not in PMD_DTYPES:
        raise TypeError("unsupported movie dtype %s" % t.dtype)
    return PMD_DTYPES[t.dtype]


# ---------------------------------------------------------------------------------------------
def stats_pass(movie2d, t_total):
    """movie2d: (t_local, d) device tensor of a supported dtype.  Returns (mean_part, noise_part, n_var):
    [n_chunks, d] float32 partials and the number of chunks that qualified for the noise estimate."""
    t_local, d = movie2d.shape
    assert movie2d.is_cuda and movie2d.is_contiguous()
    n_chunks = (t_local + 1023) // 1024
    mean_part = torch.empty((n_chunks, d), dtype=torch.float32, device=movie2d.device)
    noise_part = torch.empty((n_chunks, d), dtype=torch.float32, device=movie2d.device)
    tc, ts = _tables(movie2d.device)
    _call("pmd_stats_pass", _p(movie2d), movie_dtype_code(movie2d), t_local, d, t_total, _p(tc), _p(ts), _p(mean_part),
          _p(noise_part), _stream())
    n_var = sum(1 for c in range(n_chunks) if min(1024, t_local - 1024 * c) >= 256)
    return mean_part, noise_part, n_var


def standardize_frames(movie2d, frames, mean, stdv):
    """(movie[frames] - mean) / std as float32 (n, d).  frames: int64 device tensor."""
    _req(mean, torch.float32, "mean"), _req(stdv, torch.float32, "stdv"), _req(frames, torch.int64, "frames")
    d = movie2d.shape[1]
    n = frames.numel()
    out = torch.empty((n, d), dtype=torch.float32, device=movie2d.device)
    step = 65535
    for s in range(0, n, step):
        m = min(step, n - s)
        _call("pmd_standardize_frames", _p(movie2d), movie_dtype_code(movie2d), d, _p(frames[s:]), m, _p(mean), _p(stdv),
              _p(out[s:]), _stream())
    return out


def gram_f64(a, batch, n, m_len, batch_stride, row_stride, inner_stride):
    """Batched float64 Gram matrices of strided float32 operands (see header)."""
    _req(a, torch.float32, "a")
    c = torch.zeros((batch, n, n), dtype=torch.float64, device=a.device)
    step = 65535
    for s in range(0, batch, step):
        m = min(step, batch - s)
        _call("pmd_gram_f64", ctypes.c_void_p(a.data_ptr() + 4 * s * batch_stride), m, n, m_len, batch_stride, row_stride,
              inner_stride, _p(c[s:]), _stream())
    return c


def jacobi_eigh(c, mode=0):
    """c: (batch, n, n) float64 (destroyed).  Returns (w (batch,n) float64 descending, vecs (batch,n,n) float32)."""
    _req(c, torch.float64, "c")
    batch, n, _ = c.shape
    w = torch.empty((batch, n), dtype=torch.float64, device=c.device)
    vecs = torch.empty((batch, n, n), dtype=torch.float32, device=c.device)
    _call("pmd_jacobi_eigh", _p(c), batch, n, int(mode), _p(w), _p(vecs), _stream())
    return w, vecs


def gram_rows(x):
    """x: (batch, n, m) contiguous -> x x^T per batch."""
    b, n, m = x.shape
    return gram_f64(x, b, n, m, n * m, m, 1)


def gram_cols(x, ncols=None):
    """x: (batch, m, ld) contiguous -> x[:, :, :ncols]^T x[:, :, :ncols] per batch."""
    b, m, ld = x.shape
    n = ld if ncols is None else ncols
    return gram_f64(x, b, n, m, m * ld, 1, ld)


def orthonormalize_cols(x, ncols=None, passes=2):
    """Orthonormalise the first `ncols` columns of every (m, ld) matrix of x (batch, m, ld) by repeated
    float64-Gram / Jacobi whitening (CholQR2-like, but via the symmetric eigendecomposition so rank
    deficient inputs give zero columns instead of a breakdown).  Returns a tensor of the same shape
    with the remaining columns zero."""
    b, m, ld = x.shape
    n = ld if ncols is None else ncols
    for _ in range(passes):
        c = gram_cols(x, n)
        _, tm = jacobi_eigh(c, mode=1)
        if n != ld:
            tp = torch.zeros((b, ld, ld), dtype=torch.float32, device=x.device)
            tp[:, :n, :n] = tm
            tm = tp
        x = torch.bmm(x, tm)
    return x


def block_pool_tavg(yres, d2, starts, bh, bw, saf, taf):
    _req(yres, torch.float32, "yres"), _req(starts, torch.int32, "starts")
    t, d = yres.shape
    nb = starts.shape[0]
    ph, pw = -(-bh // saf), -(-bw // saf)
    bta = torch.empty((nb, t // taf, ph * pw), dtype=torch.float32, device=yres.device)
    _call("pmd_block_pool_tavg", _p(yres), t, d2, d, _p(starts), nb, bh, bw, saf, taf, _p(bta), _stream())
    return bta


def block_unpool(uds, bh, bw, saf, rp):
    _req(uds, torch.float32, "uds")
    nb, P, r = uds.shape
    w = torch.empty((nb, bh * bw, rp), dtype=torch.float32, device=uds.device)
    _call("pmd_block_unpool", _p(uds), nb, bh, bw, saf, r, rp, _p(w), _stream())
    return w


def block_project(movie, movie_batch_stride, t, d2, d, starts, bh, bw, w, r):
    _req(movie, torch.float32, "movie"), _req(w, torch.float32, "w"), _req(starts, torch.int32, "starts")
    nb, bpix, rp = w.shape
    assert bpix == bh * bw and starts.shape[0] == nb
    out = torch.empty((nb, r, t), dtype=torch.float32, device=movie.device)
    step = 65535
    for s in range(0, nb, step):
        m = min(step, nb - s)
        mv = ctypes.c_void_p(movie.data_ptr() + 4 * s * movie_batch_stride)
        _call("pmd_block_project", mv, movie_batch_stride, t, d2, d, _p(starts[s:]), m, bh, bw, _p(w[s:]), r, rp,
              _p(out[s:]), _stream())
    return out


def block_spatial(movie, movie_batch_stride, t, d2, d, starts, bh, bw, vb, rp):
    _req(movie, torch.float32, "movie"), _req(vb, torch.float32, "vb"), _req(starts, torch.int32, "starts")
    nb, r, tt = vb.shape
    assert tt == t
    s_out = torch.empty((nb, bh * bw, rp), dtype=torch.float32, device=movie.device)
    step = 65535
    for s in range(0, nb, step):
        m = min(step, nb - s)
        mv = ctypes.c_void_p(movie.data_ptr() + 4 * s * movie_batch_stride)
        _call("pmd_block_spatial", mv, movie_batch_stride, t, d2, d, _p(starts[s:]), m, bh, bw, _p(vb[s:]), r, rp,
              _p(s_out[s:]), _stream())
    return s_out


def block_stats_rank(u, v, bh, bw, r, thr_s, thr_t, max_fail):
    _req(u, torch.float32, "u"), _req(v, torch.float32, "v")
    nb, bpix, rp = u.shape
    t = v.shape[2]
    sstat = torch.empty((nb, r), dtype=torch.float32, device=u.device)
    tstat = torch.empty((nb, r), dtype=torch.float32, device=u.device)
    ranks = torch.empty((nb,), dtype=torch.int32, device=u.device)
    _call("pmd_block_stats_rank", _p(u), _p(v), nb, bh, bw, r, rp, t, float(thr_s), float(thr_t), int(max_fail), _p(sstat),
          _p(tstat), _p(ranks), _stream())
    return sstat, tstat, ranks


def assemble_u(u, bh, bw, starts, ranks, col0, block_weights, cumw, d2, n_cols):
    _req(u, torch.float32, "u"), _req(ranks, torch.int32, "ranks"), _req(col0, torch.int64, "col0")
    _req(block_weights, torch.float32, "block_weights"), _req(cumw, torch.float64, "cumw")
    nb, bpix, rp = u.shape
    uv64 = torch.empty((n_cols, bpix), dtype=torch.float64, device=u.device)
    uv32 = torch.empty((n_cols, bpix), dtype=torch.float32, device=u.device)
    _call("pmd_assemble_u", _p(u), nb, bh, bw, rp, _p(starts), _p(ranks), _p(col0), _p(block_weights), _p(cumw), d2,
          _p(uv64), _p(uv32), _stream())
    return uv64, uv32


def make_tasks(ranks_host):
    """(block, first component) for every group of <= 4 kept components (host logic for pmd_project_local)."""
    ranks_host = np.asarray(ranks_host, dtype=np.int64)
    ngrp = (ranks_host + 3) // 4
    blk = np.repeat(np.arange(len(ranks_host)), ngrp)
    first = np.concatenate([np.arange(g) * 4 for g in ngrp]) if len(ranks_host) else np.zeros(0, dtype=np.int64)
    return np.stack([blk, first], axis=1).astype(np.int32)


def project_local(movie2d, d2, starts, bh, bw, ranks, col0, tasks, uvals32, mean, inv_std, z):
    """z[col, f] (+)= U_loc^T standardised movie, z: (n_cols, ldz) float32 view whose first movie2d.shape[0]
    columns are written."""
    t, d = movie2d.shape
    _req(uvals32, torch.float32, "uvals32"), _req(tasks, torch.int32, "tasks")
    assert z.dtype == torch.float32 and z.stride(1) == 1
    _call("pmd_project_local", _p(movie2d), movie_dtype_code(movie2d), t, d2, d, _p(starts), starts.shape[0], bh, bw,
          _p(ranks), _p(col0), _p(tasks), tasks.shape[0], _p(uvals32), _p(mean), _p(inv_std), _p(z), z.stride(0), _stream())


def project_dense(movie2d, basis, mean, inv_std, z):
    """z[c, f] += basis[c] . standardised frame f; z: (k, ldz) float32 view, zero-initialised by the caller."""
    t, d = movie2d.shape
    _req(basis, torch.float32, "basis")
    k = basis.shape[0]
    assert z.dtype == torch.float32 and z.stride(1) == 1
    for c0 in range(0, k, 16):
        kk = min(16, k - c0)
        _call("pmd_project_dense", _p(movie2d), movie_dtype_code(movie2d), t, d, _p(basis[c0:]), kk, _p(mean), _p(inv_std),
              _p(z[c0:]), z.stride(0), _stream())


def reconstruct(indptr, indices, values, c, pix, scale, shift):
    """out[n, i] = (U[pix[i], :] @ c[:, n]) * scale[pix[i]] + shift[pix[i]]  ->  (n, npix) float32."""
    _req(indptr, torch.int64, "indptr"), _req(indices, torch.int32, "indices"), _req(values, torch.float32, "values")
    _req(c, torch.float32, "c"), _req(pix, torch.int32, "pix")
    n = c.shape[1]
    npix = pix.numel()
    out = torch.empty((n, npix), dtype=torch.float32, device=c.device)
    _call("pmd_reconstruct", _p(indptr), _p(indices), _p(values), _p(c), n, _p(pix), npix, _p(scale), _p(shift), _p(out),
          _stream())
    return out


def reconstruct_f64(indptr, indices, values64, c64, pix):
    """(n, npix) float64 = (U[pix] @ c)^T with float64 CSR values and coefficients (whitening step)."""
    _req(indptr, torch.int64, "indptr"), _req(indices, torch.int32, "indices"), _req(values64, torch.float64, "values")
    _req(c64, torch.float64, "c"), _req(pix, torch.int32, "pix")
    n = c64.shape[1]
    out = torch.empty((n, pix.numel()), dtype=torch.float64, device=c64.device)
    _call("pmd_reconstruct_f64", _p(indptr), _p(indices), _p(values64), _p(c64), n, _p(pix), pix.numel(), _p(out), _stream())
    return out


def project_cols_f64(w64, d2, starts, bh, bw, blk_of_col, col0, uvals64, bg64):
    """Z (n_cols, m) float64 = U^T w64^T for w64 (m, d) float64."""
    _req(w64, torch.float64, "w"), _req(uvals64, torch.float64, "uvals64"), _req(bg64, torch.float64, "bg64")
    _req(blk_of_col, torch.int32, "blk_of_col")
    m, d = w64.shape
    n_local = uvals64.shape[0]
    n_cols = n_local + bg64.shape[0]
    z = torch.empty((n_cols, m), dtype=torch.float64, device=w64.device)
    _call("pmd_project_cols_f64", _p(w64), m, d2, d, _p(starts), bh, bw, _p(blk_of_col), _p(col0), n_local, _p(uvals64),
          _p(bg64), n_cols, _p(z), _stream())
    return z


def split_groups(rk):
    """Split `rk` kept components into ceil(rk/4) groups of nearly equal size (each <= 4)."""
    ng = (rk + 3) // 4
    base, extra = divmod(rk, ng)
    return [base + (1 if i < extra else 0) for i in range(ng)]


def make_supertiles(row_starts, col_starts, bh, bw, ranks_host, col0_host, max_pixels=2048, target_tasks=14):
    """Host tables for pmd_project_supertile: group G x G neighbouring blocks (block grid = row_starts x col_starts,
    blocks numbered row-major) so that the staged pixel region stays <= max_pixels and the mean number of
    (block, component group) tasks per supertile is about `target_tasks` (one task per warp).
    Returns dict(tiles int32 [n,4], task_ptr int32 [n+1], tasks int32 [m,4], max_h, max_w, G)."""
    row_starts, col_starts = list(row_starts), list(col_starts)
    nbr, nbc = len(row_starts), len(col_starts)
    ranks_host = np.asarray(ranks_host, dtype=np.int64).reshape(nbr, nbc)
    col0_host = np.asarray(col0_host, dtype=np.int64).reshape(nbr, nbc)
    groups_per_block = float(np.mean((ranks_host + 3) // 4))

    def region(g):
        h = max(row_starts[min(a + g, nbr) - 1] + bh - row_starts[a] for a in range(0, nbr, g))
        w = max(col_starts[min(c + g, nbc) - 1] + bw - col_starts[c] for c in range(0, nbc, g))
        return h, w

    G = 1
    for g in range(2, 9):
        h, w = region(g)
        if h * w > max_pixels or g * g * groups_per_block > target_tasks * 1.15:
            break
        G = g
    tiles, task_ptr, tasks = [], [0], []
    for a0 in range(0, nbr, G):
        for c0 in range(0, nbc, G):
            a1, c1 = min(a0 + G, nbr), min(c0 + G, nbc)
            r0, cc0 = row_starts[a0], col_starts[c0]
            tiles.append((r0, cc0, row_starts[a1 - 1] + bh - r0, col_starts[c1 - 1] + bw - cc0))
            for a in range(a0, a1):
                for c in range(c0, c1):
                    first = int(col0_host[a, c])
                    for n in split_groups(int(ranks_host[a, c])):
                        tasks.append((row_starts[a] - r0, col_starts[c] - cc0, first, n))
                        first += n
            task_ptr.append(len(tasks))
    tiles = np.array(tiles, dtype=np.int32).reshape(-1, 4)
    return dict(tiles=tiles, task_ptr=np.array(task_ptr, dtype=np.int32), tasks=np.array(tasks, dtype=np.int32).reshape(-1, 4),
                max_h=int(tiles[:, 2].max()), max_w=int(tiles[:, 3].max()), G=G)


def project_supertile(movie2d, d2, st, bh, bw, uvals32, mean, inv_std, z):
    """z[col, f] = U_loc^T standardised movie via the supertile kernel; `st` = device copies of make_supertiles()."""
    t, d = movie2d.shape
    _req(uvals32, torch.float32, "uvals32")
    assert z.dtype == torch.float32 and z.stride(1) == 1
    _call("pmd_project_supertile", _p(movie2d), movie_dtype_code(movie2d), t, d2, d, _p(st["tiles"]), st["tiles"].shape[0],
          _p(st["task_ptr"]), _p(st["tasks"]), bh, bw, st["max_h"], st["max_w"], _p(uvals32), _p(mean), _p(inv_std), _p(z),
          z.stride(0), _stream())
