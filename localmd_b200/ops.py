"""Thin torch-tensor wrappers around the C ABI (include/pmd_sm100.h).

PyTorch is plumbing only here: it owns device memory and streams; every wrapper validates shapes /
dtypes / contiguity, allocates outputs and enqueues one kernel (or a short fixed sequence) on the
current CUDA stream through ctypes.  No wrapper has a CPU path."""
import contextlib
import os
import ctypes

import numpy as np
import torch

from . import _lib
from ._tables import welch_fft_tables, welch_tc_tables

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
_KERNELS_PER_CALL = {"pmd_block_stats_rank": 3, "pmd_block_pool_full": 2, "pmd_sym_product_f64": 2}


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
        _table_cache[key] = torch.from_numpy(welch_fft_tables()).to(device)
    return _table_cache[key]


def _tables_tc(device):
    key = "tc:" + str(device)
    if key not in _table_cache:
        _table_cache[key] = torch.from_numpy(welch_tc_tables()).to(device)
    return _table_cache[key]


def h2d(a, device, dtype=None):
    """Host array -> device tensor (small tables).  Deliberately the plain synchronous copy: staging these uploads through
    page-locked memory with non_blocking copies (so that the host can enqueue further ahead of the device) was measured
    at C2 and made the job SLOWER whenever stage timings were requested (host stalls of 2-9 ms between consecutive table
    uploads in the whitening / projection stages, 107 -> 121 ms per job); see DESIGN.md section 5."""
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device)


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
    which = os.environ.get("PMD_K1", "tc")   # development switch between the generations of the statistics kernel
    if which == "tc" and movie2d.data_ptr() % 16 == 0 and (d * movie2d.element_size()) % 16 == 0:
        # K1 on the tensor cores (csrc/stats_tc.cu): 2-D TMA boxes need 16-byte aligned frames
        _call("pmd_stats_pass_tc", _p(movie2d), movie_dtype_code(movie2d), t_local, d, t_total, _p(_tables_tc(movie2d.device)),
              _p(mean_part), _p(noise_part), _stream())
    else:
        _call("pmd_stats_pass", _p(movie2d), movie_dtype_code(movie2d), t_local, d, t_total, _p(_tables(movie2d.device)),
              _p(mean_part), _p(noise_part), _stream())
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



_SYM_TILE = 128
_sym_work = {}


def sym_splits(n, k_len, sms=148):
    """Number of inner-dimension chunks of pmd_sym_product_f64: the split whose nt (nt + 1) / 2 * splits units fill whole
    waves of the SMs best (one 128 x 128 unit per SM at a time), chunks of at least 256 values, fewest chunks on ties."""
    nt = -(-n // _SYM_TILE)
    tiles = nt * (nt + 1) // 2
    best, best_eff = 1, 0.0
    for s in range(1, 65):
        if s > 1 and k_len // s < 256:
            break
        units = tiles * s
        eff = units / (sms * -(-units // sms))
        if eff > best_eff + 1e-9:
            best, best_eff = s, eff
    return best


def sym_product_f64(a, b=None, layout=0):
    """float64 product known to be SYMMETRIC, on the FP64 tensor cores (csrc/sym_f64.cu; see header):
    layout 0: a (n, k) [, b (n, k)] -> a b^T (b None: the Gram a a^T);  layout 1: a (k, n), b (k, n) -> a^T b.
    Operands float32 or float64 (layout 0: both the same type; layout 1: b float64), rows contiguous."""
    if a.dim() != 2 or a.stride(1) != 1 or not a.is_cuda:
        raise ValueError("a must be a 2-D CUDA tensor with contiguous rows (there is no CPU fallback)")
    if b is not None and (b.dim() != 2 or b.stride(1) != 1 or b.shape != a.shape or b.device != a.device):
        raise ValueError("b must match a")
    if a.dtype not in (torch.float32, torch.float64) or (b is not None and b.dtype not in (torch.float32, torch.float64)):
        raise TypeError("operands must be float32 or float64")
    n, k_len = (a.shape[0], a.shape[1]) if layout == 0 else (a.shape[1], a.shape[0])
    splits = sym_splits(n, k_len)
    nt = -(-n // _SYM_TILE)
    need = splits * (nt * (nt + 1) // 2) * _SYM_TILE * _SYM_TILE
    key = (str(a.device), torch.cuda.current_stream().cuda_stream)
    work = _sym_work.get(key)
    if work is None or work.numel() < need:
        work = _sym_work[key] = torch.empty(need, dtype=torch.float64, device=a.device)
    c = torch.empty((n, n), dtype=torch.float64, device=a.device)
    _call("pmd_sym_product_f64", _p(a), PMD_DTYPES[a.dtype], a.stride(0), _p(b), PMD_DTYPES[b.dtype] if b is not None else 0,
          b.stride(0) if b is not None else 0, int(layout), n, k_len, splits, _p(work), _p(c), _stream())
    return c


def jacobi_eigh(c, mode=0, sweeps_f32=False):
    """c: (batch, n, n) float64 (destroyed).  Returns (w (batch,n) float64 descending, vecs (batch,n,n) float32).
    sweeps_f32: run the Jacobi rotations in float32 (reference-level accuracy; ~10x faster on this GPU)."""
    _req(c, torch.float64, "c")
    batch, n, _ = c.shape
    w = torch.empty((batch, n), dtype=torch.float64, device=c.device)
    vecs = torch.empty((batch, n, n), dtype=torch.float32, device=c.device)
    _call("pmd_jacobi_eigh", _p(c), batch, n, int(mode), int(bool(sweeps_f32)), _p(w), _p(vecs), _stream())
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


def standardize_frames_t(movie2d, frames, mean, stdv, ld=None):
    """Pixel-major standardised init movie: out[p, i] = (movie[frames[i], p] - mean[p]) / std[p] as float32
    (d, ld); ld defaults to len(frames) rounded up to a multiple of 4, padding columns are zero."""
    _req(mean, torch.float32, "mean"), _req(stdv, torch.float32, "stdv"), _req(frames, torch.int64, "frames")
    d = movie2d.shape[1]
    n = frames.numel()
    ld = (n + 3) // 4 * 4 if ld is None else int(ld)
    out = torch.empty((d, ld), dtype=torch.float32, device=movie2d.device)
    _call("pmd_standardize_frames_t", _p(movie2d), movie_dtype_code(movie2d), d, _p(frames), n, _p(mean), _p(stdv), _p(out), ld,
          _stream())
    return out


def rows_sketch(yt, n, omega):
    """y (d, l) = yt[:, :n] @ omega for the pixel-major standardised frames yt (d, ld) and a sketch matrix omega (n, l),
    l <= 32 (pmd_loader.py:58, the random projection of the background rSVD)."""
    _req(yt, torch.float32, "yt"), _req(omega, torch.float32, "omega")
    d, ld = yt.shape
    l = omega.shape[1]
    y = torch.empty((d, l), dtype=torch.float32, device=yt.device)
    _call("pmd_rows_sketch", _p(yt), ld, d, int(n), _p(omega), l, _p(y), l, _stream())
    return y


def chol_whiten(g):
    """(batch, n, n) float64 Gram matrices -> (batch, n, n) float32 L^-T of their Cholesky factors, n <= 32."""
    _req(g, torch.float64, "g")
    b, n, _ = g.shape
    t = torch.empty((b, n, n), dtype=torch.float32, device=g.device)
    _call("pmd_chol_whiten", _p(g), b, n, _p(t), _stream())
    return t


def rows_times_small(x, m, ncols=None, transposed=False):
    """x (batch, d, ldx)[:, :, :k] @ m (batch, k, nc) with k, nc <= 32 -> (batch, d, nc), or (batch, nc, d) when
    `transposed`.  k = m.shape[1]."""
    _req(x, torch.float32, "x"), _req(m, torch.float32, "m")
    b, d, ldx = x.shape
    k, nc = m.shape[1], (m.shape[2] if ncols is None else int(ncols))
    out = torch.empty((b, nc, d) if transposed else (b, d, nc), dtype=torch.float32, device=x.device)
    _call("pmd_rows_times_small", _p(x), ldx, d, k, _p(m), m.shape[2], nc, _p(out), d if transposed else nc, int(bool(transposed)),
          b, d * ldx, k * m.shape[2], d * nc, _stream())
    return out


def bg_project_t(yt, bg, n_ranges=128):
    """vbg (K, ld) = bg (K, d) @ yt (d, ld) for the pixel-major init movie, K <= 32 (deterministic two-stage sum)."""
    _req(yt, torch.float32, "yt"), _req(bg, torch.float32, "bg")
    d, ld = yt.shape
    k = bg.shape[0]
    n_ranges = max(1, min(int(n_ranges), (d + 255) // 256))
    part = torch.empty((n_ranges, k, ld), dtype=torch.float32, device=yt.device)
    _call("pmd_bg_project_t", _p(yt), ld, d, _p(bg), k, n_ranges, _p(part), _stream())
    return part.sum(dim=0)


def bg_remove_t(yt, bg, vbg):
    """yt (d, ld) -= bg^T (d, K) @ vbg (K, ld), in place, K <= 16."""
    _req(yt, torch.float32, "yt"), _req(bg, torch.float32, "bg"), _req(vbg, torch.float32, "vbg")
    d, ld = yt.shape
    _call("pmd_bg_remove_t", _p(yt), ld, d, _p(bg), bg.shape[0], _p(vbg), _stream())
    return yt


def _split_tf32(x):
    """x = hi + lo with hi exactly representable in TF32 (10 mantissa bits, round to nearest)."""
    hi = ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, x - hi


@contextlib.contextmanager
def fp32_matmul():
    """Library GEMMs inside this block run in full float32 whatever the host application set globally
    (torch.backends.cuda.matmul.allow_tf32 / set_float32_matmul_precision): the parity bars depend on it."""
    old_tf32, old_prec = torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        yield
    finally:
        torch.set_float32_matmul_precision(old_prec)
        torch.backends.cuda.matmul.allow_tf32 = old_tf32


def matmul_3xtf32(a, b):
    """float32-accurate a @ b on the tensor cores: three TF32 library GEMMs on split operands
    (a_hi b_hi + a_lo b_hi + a_hi b_lo; the dropped a_lo b_lo term is ~2^-22 relative).  Used for the mixing GEMM
    P^T (U^T Y) of pmd_loader.py:411-412."""
    a_hi, a_lo = _split_tf32(a.contiguous())
    b_hi, b_lo = _split_tf32(b.contiguous())
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = torch.matmul(a_hi, b_lo)
        out.addmm_(a_lo, b_hi)
        out.addmm_(a_hi, b_hi)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def split_pairs(x, mode):
    """In place x <- hi(x) (TF32-exact part); returns the bf16 correction operand: mode 0 (x is the right operand, (K, n)) ->
    (2K, n) = [bf16(hi); bf16(lo)], mode 1 (x is the left operand, (m, K)) -> (m, 2K) = [bf16(lo) | bf16(hi)]."""
    _req(x, torch.float32, "x")
    rows, cols = x.shape
    pair = torch.empty((2 * rows, cols) if mode == 0 else (rows, 2 * cols), dtype=torch.bfloat16, device=x.device)
    _call("pmd_split_tf32_bf16", _p(x), rows, cols, cols, _p(pair), pair.shape[1], int(mode), _stream())
    return pair


def matmul_split(a_hi, a_pair, b_hi, b_pair):
    """a @ b at float32 accuracy from operands prepared by split_pairs: one TF32 library GEMM (exact products of the hi
    parts) + one bf16 library GEMM of twice the depth carrying a_lo b_hi + a_hi b_lo (dropped terms < 2^-18 relative)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = torch.mm(a_pair, b_pair, out_dtype=torch.float32)
        out.addmm_(a_hi, b_hi)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def matmul_3xtf32_any(a, b):
    """float32-accurate a @ b on the tensor cores for arbitrary shapes: every dimension is zero padded to a multiple of 8
    (the tensor-core kernels want 16-byte aligned rows) into fresh copies, which are split in place (split_pairs) and
    multiplied by matmul_split (one TF32 + one bf16 library GEMM); the result is sliced back."""
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    mp, kp, np_ = (m + 7) // 8 * 8, (k + 7) // 8 * 8, (n + 7) // 8 * 8
    if (mp, kp) != (m, k) or not a.is_contiguous():
        ap = torch.zeros((mp, kp), dtype=torch.float32, device=a.device)
        ap[:m, :k] = a
    else:
        ap = a.clone()
    if (kp, np_) != (k, n) or not b.is_contiguous():
        bp = torch.zeros((kp, np_), dtype=torch.float32, device=b.device)
        bp[:k, :n] = b
    else:
        bp = b.clone()
    return matmul_split(ap, split_pairs(ap, 1), bp, split_pairs(bp, 0))[:m, :n]


def block_orth_fits(m, n):
    """True when an (m, n) matrix (+ its float64 Gram) fits the shared memory of pmd_block_orth."""
    n4, n8, m8 = (n + 3) // 4 * 4, (n + 7) // 8 * 8, (m + 7) // 8 * 8
    ldg = n8 if n8 & 8 else n8 + 8
    return n <= 64 and (n8 * ldg + 3 * n8) * 8 + m8 * n4 * 4 <= 227 * 1024


def block_orth(x, ncols=None, g_ext=None, passes=2):
    """In-place fused orthonormalisation (pmd_block_orth) of the first ncols columns of every (m, ld) matrix of
    x (batch, m, ld); optional float64 Gram g_ext (batch, ncols, ncols) for the pre-whitening solve.  Returns x."""
    _req(x, torch.float32, "x")
    b, m, ld = x.shape
    n = ld if ncols is None else int(ncols)
    if g_ext is not None:
        _req(g_ext, torch.float64, "g_ext")
        assert tuple(g_ext.shape) == (b, n, n)
    _call("pmd_block_orth", _p(x), b, m, n, ld, _p(g_ext), passes, _stream())
    return x


def block_pool_tavg(yt, t, d2, starts, bh, bw, saf, taf):
    """(nb, P, t // taf) pooled + time-averaged blocks of the pixel-major init movie yt (d, ld)."""
    _req(yt, torch.float32, "yt"), _req(starts, torch.int32, "starts")
    d, ld = yt.shape
    nb = starts.shape[0]
    ph, pw = -(-bh // saf), -(-bw // saf)
    bta = torch.empty((nb, ph * pw, t // taf), dtype=torch.float32, device=yt.device)
    _call("pmd_block_pool_tavg", _p(yt), ld, t, d2, _p(starts), nb, bh, bw, saf, taf, _p(bta), _stream())
    return bta


def block_pool_full(yt, t, d2, starts, bh, bw, saf, taf):
    """(pooled (nb, P, ld) full time resolution, bta (nb, P, t // taf)) of the pixel-major init movie yt (d, ld)."""
    _req(yt, torch.float32, "yt"), _req(starts, torch.int32, "starts")
    d, ld = yt.shape
    nb = starts.shape[0]
    ph, pw = -(-bh // saf), -(-bw // saf)
    pooled = torch.empty((nb, ph * pw, ld), dtype=torch.float32, device=yt.device)
    bta = torch.empty((nb, ph * pw, t // taf), dtype=torch.float32, device=yt.device)
    _call("pmd_block_pool_full", _p(yt), ld, t, d2, _p(starts), nb, bh, bw, saf, taf, _p(pooled), _p(bta), _stream())
    return pooled, bta


def block_unpool(uds, bh, bw, saf, rp):
    _req(uds, torch.float32, "uds")
    nb, P, r = uds.shape
    w = torch.empty((nb, bh * bw, rp), dtype=torch.float32, device=uds.device)
    _call("pmd_block_unpool", _p(uds), nb, bh, bw, saf, r, rp, _p(w), _stream())
    return w


def block_project(movie_t, movie_batch_stride, ld, d2, starts, bh, bw, w, r, ldo=None):
    """out[b, c, f] = sum_q w[b, q, c] * Y_b[q, f] over the pixel-major movie (leading dimension ld).
    Returns (nb, r, ldo) float32 (ldo defaults to ld; columns past the data are zero)."""
    _req(movie_t, torch.float32, "movie_t"), _req(w, torch.float32, "w"), _req(starts, torch.int32, "starts")
    nb, bpix, rp = w.shape
    assert bpix == bh * bw and starts.shape[0] == nb
    ldo = ld if ldo is None else int(ldo)
    out = torch.empty((nb, r, ldo), dtype=torch.float32, device=movie_t.device)
    step = 65535
    for s in range(0, nb, step):
        m = min(step, nb - s)
        mv = ctypes.c_void_p(movie_t.data_ptr() + 4 * s * movie_batch_stride)
        _call("pmd_block_project", mv, movie_batch_stride, ld, d2, _p(starts[s:]), m, bh, bw, _p(w[s:]), r, rp, _p(out[s:]),
              ldo, _stream())
    return out


def block_project_tc(movie_t, movie_batch_stride, ld, d2, starts, bh, bw, w, r, ldo=None):
    """block_project on the tcgen05 tensor cores (3xTF32 with float32 accumulation in tensor memory).  The kernels hold at
    most 64 components per accumulator tile; wider coefficient sets (max_components > 64) go through in slices of 64."""
    _req(movie_t, torch.float32, "movie_t"), _req(w, torch.float32, "w"), _req(starts, torch.int32, "starts")
    nb, bpix, rp = w.shape
    assert bpix == bh * bw and starts.shape[0] == nb
    ldo = ld if ldo is None else int(ldo)
    if rp > 64:
        outs = []
        for c0 in range(0, r, 64):
            rc = min(64, r - c0)
            wc = torch.zeros((nb, bpix, (rc + 3) // 4 * 4), dtype=torch.float32, device=w.device)
            wc[:, :, :rc] = w[:, :, c0 : c0 + rc]
            outs.append(block_project_tc(movie_t, movie_batch_stride, ld, d2, starts, bh, bw, wc, rc, ldo))
        return torch.cat(outs, dim=1)
    out = torch.empty((nb, r, ldo), dtype=torch.float32, device=movie_t.device)
    which = os.environ.get("PMD_BLOCK_PROJECT", "ts")   # development switch between the generations of the kernel
    n_rows = movie_t.numel() // ld
    if (which == "ts" and bw % 2 == 0 and movie_batch_stride % ld == 0 and movie_t.data_ptr() % 16 == 0 and movie_t.is_contiguous()
            and n_rows < 2**31):
        # movie operand in tensor memory (csrc/blocks_ts.cu)
        wc = w.contiguous()
        step = 65535
        for s in range(0, nb, step):
            m = min(step, nb - s)
            ws = torch.empty(int(_lib.lib().pmd_block_project_ts_workspace_bytes(m, bh, bw)), dtype=torch.uint8, device=movie_t.device)
            mv = ctypes.c_void_p(movie_t.data_ptr() + 4 * s * movie_batch_stride)
            _call("pmd_block_project_ts", mv, movie_batch_stride, n_rows - s * (movie_batch_stride // ld), ld, d2, _p(starts[s:]), m, bh, bw,
                  _p(wc[s:]), r, rp, _p(ws), _p(out[s:]), ldo, _stream())
        return out
    w_hi, w_lo = _split_tf32(w)
    step = 65535
    for s in range(0, nb, step):
        m = min(step, nb - s)
        mv = ctypes.c_void_p(movie_t.data_ptr() + 4 * s * movie_batch_stride)
        _call("pmd_block_project_tc", mv, movie_batch_stride, ld, d2, _p(starts[s:]), m, bh, bw, _p(w_hi[s:]), _p(w_lo[s:]), r, rp,
              _p(out[s:]), ldo, _stream())
    return out


def block_spatial(movie_t, movie_batch_stride, ld, d2, starts, bh, bw, v, rp):
    """s[b, q, c] = sum_f Y_b[q, f] * v[b, c, f];  v (nb, r, ldv) with zero padding -> (nb, bh*bw, rp)."""
    _req(movie_t, torch.float32, "movie_t"), _req(v, torch.float32, "v"), _req(starts, torch.int32, "starts")
    nb, r, ldv = v.shape
    s_out = torch.empty((nb, bh * bw, rp), dtype=torch.float32, device=movie_t.device)
    step = 65535
    for s in range(0, nb, step):
        m = min(step, nb - s)
        mv = ctypes.c_void_p(movie_t.data_ptr() + 4 * s * movie_batch_stride)
        _call("pmd_block_spatial", mv, movie_batch_stride, ld, d2, _p(starts[s:]), m, bh, bw, _p(v[s:]), ldv, r, rp,
              _p(s_out[s:]), _stream())
    return s_out


def block_spatial_tc(movie_t, movie_batch_stride, ld, d2, starts, bh, bw, v, rp):
    """block_spatial on the tcgen05 tensor cores (3xTF32, float32 accumulation in tensor memory); more than 64 components
    (max_components > 64) go through in slices of 64."""
    _req(movie_t, torch.float32, "movie_t"), _req(v, torch.float32, "v"), _req(starts, torch.int32, "starts")
    nb, r, ldv = v.shape
    if rp > 64:
        s_out = torch.zeros((nb, bh * bw, rp), dtype=torch.float32, device=movie_t.device)
        for c0 in range(0, r, 64):
            rc = min(64, r - c0)
            s_out[:, :, c0 : c0 + rc] = block_spatial_tc(movie_t, movie_batch_stride, ld, d2, starts, bh, bw,
                                                         v[:, c0 : c0 + rc].contiguous(), (rc + 3) // 4 * 4)[:, :, :rc]
        return s_out
    s_out = torch.empty((nb, bh * bw, rp), dtype=torch.float32, device=movie_t.device)
    if os.environ.get("PMD_BLOCK_SPATIAL", "ts") == "ts":   # development switch between the generations of the kernel
        _call("pmd_block_spatial_ts", _p(movie_t), movie_batch_stride, ld, d2, _p(starts), nb, bh, bw, _p(v.contiguous()), ldv, r, rp,
              _p(s_out), _stream())
        return s_out
    step = 65535
    for s in range(0, nb, step):
        m = min(step, nb - s)
        mv = ctypes.c_void_p(movie_t.data_ptr() + 4 * s * movie_batch_stride)
        _call("pmd_block_spatial_tc", mv, movie_batch_stride, ld, d2, _p(starts[s:]), m, bh, bw, _p(v[s:]), ldv, r, rp,
              _p(s_out[s:]), _stream())
    return s_out


def block_stats_rank(u, v, bh, bw, r, thr_s, thr_t, max_fail, t=None):
    """v: (nb, r, ldv); the temporal statistic uses the first t columns of every row (default: all)."""
    _req(u, torch.float32, "u"), _req(v, torch.float32, "v")
    nb, bpix, rp = u.shape
    ldv = v.shape[2]
    t = ldv if t is None else int(t)
    sstat = torch.empty((nb, r), dtype=torch.float32, device=u.device)
    tstat = torch.empty((nb, r), dtype=torch.float32, device=u.device)
    ranks = torch.empty((nb,), dtype=torch.int32, device=u.device)
    _call("pmd_block_stats_rank", _p(u), _p(v), nb, bh, bw, r, rp, t, ldv, float(thr_s), float(thr_t), int(max_fail), _p(sstat),
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
    # the kernel gives one warp to every column: right for the block-supported columns (b pixels each), but a dense
    # background column walks all d pixels -- those K rows are a plain (K x d)(d x m) library GEMM instead
    if n_local > 0:
        _call("pmd_project_cols_f64", _p(w64), m, d2, d, _p(starts), bh, bw, _p(blk_of_col), _p(col0), n_local, _p(uvals64),
              _p(bg64), n_local, _p(z), _stream())
    if n_cols > n_local:
        torch.matmul(bg64, w64.t(), out=z[n_local:])
    return z


_pairs_cache = {}


def overlap_pairs(starts, bh, bw):
    """All ordered pairs (b1, b2) of blocks whose windows intersect, sorted by (b1, b2).  Host logic; the result
    depends on the block grid only and is memoised (a warm process decomposes many movies of one geometry)."""
    starts = np.asarray(starts, dtype=np.int64).reshape(-1, 2)
    key = (int(bh), int(bw), starts.tobytes())
    hit = _pairs_cache.get(key)
    if hit is None:
        if len(_pairs_cache) > 8:
            _pairs_cache.clear()
        hit = _pairs_cache[key] = _overlap_pairs(starts, bh, bw)
    return hit


def _overlap_pairs(starts, bh, bw):
    nb = starts.shape[0]
    rows, cols = np.unique(starts[:, 0]), np.unique(starts[:, 1])
    if len(rows) * len(cols) == nb and np.array_equal(
            starts, np.stack(np.meshgrid(rows, cols, indexing="ij"), axis=-1).reshape(-1, 2)):
        ar, ar2 = np.nonzero(np.abs(rows[:, None] - rows[None, :]) < bh)
        ac, ac2 = np.nonzero(np.abs(cols[:, None] - cols[None, :]) < bw)
        nbc = len(cols)
        b1 = (ar[:, None] * nbc + ac[None, :]).reshape(-1)
        b2 = (ar2[:, None] * nbc + ac2[None, :]).reshape(-1)
    else:  # arbitrary block lists
        hit = (np.abs(starts[:, None, 0] - starts[None, :, 0]) < bh) & (np.abs(starts[:, None, 1] - starts[None, :, 1]) < bw)
        b1, b2 = np.nonzero(hit)
    order = np.lexsort((b2, b1))
    return np.stack([b1[order], b2[order]], axis=1).astype(np.int32)


def utu_host_tables(starts_host, bh, bw, ranks_host):
    """Host bookkeeping of U_loc^T U_loc: (pairs int32 [n, 2], rowoff int64 [n], rowptr int64 [n_local + 1]); depends on the
    block grid and the kept ranks only.  The loops run in the native library (pmd_utu_host_tables: the ctypes call
    releases the interpreter lock, so the driver's worker thread does not stall the launching thread)."""
    ranks_host = np.ascontiguousarray(ranks_host, dtype=np.int64)
    pairs = overlap_pairs(starts_host, bh, bw)
    rowoff = np.empty(pairs.shape[0], dtype=np.int64)
    rowptr = np.empty(int(ranks_host.sum()) + 1, dtype=np.int64)
    _lib.check(_lib.lib().pmd_utu_host_tables(_np_ptr(pairs), pairs.shape[0], _np_ptr(ranks_host), len(ranks_host), _np_ptr(rowoff),
                                              _np_ptr(rowptr)), "pmd_utu_host_tables")
    return pairs, rowoff, rowptr


def utu_local_csr(starts_host, starts, bh, bw, ranks_host, ranks, col0_host, col0, uvals64, host=None):
    """Canonical CSR (rowptr int64, cols int32, vals float64) of U_loc^T U_loc on the device, plus the tile tables
    (pairs, seg_ptr, pair_rowoff) that utu_apply_tiles walks.  host: the result of utu_host_tables for the same
    arguments when it was computed ahead of time."""
    _req(uvals64, torch.float64, "uvals64"), _req(ranks, torch.int32, "ranks"), _req(col0, torch.int64, "col0")
    dev = uvals64.device
    pairs, rowoff, rowptr = host if host is not None else utu_host_tables(starts_host, bh, bw, ranks_host)
    nnz = int(rowptr[-1])
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    cols = torch.empty(nnz, dtype=torch.int32, device=dev)
    rowptr_d = h2d(rowptr, dev)
    tiles = None
    if nnz:
        pairs_d = h2d(pairs, dev)  # named: the tensors must outlive the pointer extraction
        rowoff_d = h2d(rowoff, dev)
        _call("pmd_utu_pairs", _p(pairs_d), pairs.shape[0], _p(rowoff_d), _p(starts), bh, bw, _p(ranks), _p(col0), _p(uvals64),
              _p(rowptr_d), _p(vals), _p(cols), _stream())
        seg = np.searchsorted(pairs[:, 0], np.arange(len(ranks_host) + 1), side="left").astype(np.int32)   # pairs are sorted by b1
        tiles = (pairs_d, h2d(seg, dev), rowoff_d)
    return rowptr_d, cols, vals, tiles


def utu_apply_tiles(tiles, ranks, col0, rowptr, vals, x, z):
    """z[:n_local] = (U_loc^T U_loc) x in float64 from the dense block-pair tiles (pmd_utu_apply_tiles, FP64 tensor cores).
    x (n_local, m), z (n_local, m): float64, rows contiguous."""
    _req(vals, torch.float64, "vals"), _req(ranks, torch.int32, "ranks"), _req(col0, torch.int64, "col0")
    if x.dtype != torch.float64 or z.dtype != torch.float64 or x.stride(1) != 1 or z.stride(1) != 1 or not x.is_cuda:
        raise ValueError("x and z must be float64 CUDA tensors with contiguous rows")
    pairs_d, seg_d, rowoff_d = tiles
    _call("pmd_utu_apply_tiles", _p(pairs_d), _p(seg_d), ranks.numel(), _p(rowoff_d), _p(ranks), _p(col0), _p(rowptr), _p(vals),
          _p(x), x.stride(0), x.shape[1], _p(z), z.stride(0), _stream())
    return z


def export_csr(uvals64, bg, d1, d2, row_starts, col_starts, bh, bw, ranks, col0, n_local, row_ids):
    """Canonical CSR of U straight from the block-component form (pmd_export_csr; see header): returns
    ((indptr_rel, cols_rel, vals_rel64), (indptr_phys, cols_phys, vals_phys32)) with the cols / vals buffers sized for the
    worst case (every stored value nonzero); the true number of entries is indptr[-1] (the caller slices once it reads it)."""
    _req(uvals64, torch.float64, "uvals64"), _req(bg, torch.float32, "bg"), _req(ranks, torch.int32, "ranks"), _req(col0, torch.int64, "col0")
    dev = uvals64.device
    d, K = d1 * d2, bg.shape[0]
    rs, cs = row_starts, col_starts     # int32 device tensors (ascending block-row / block-column origins)
    _req(rs, torch.int32, "row_starts"), _req(cs, torch.int32, "col_starts")
    if row_ids is not None:
        _req(row_ids, torch.int64, "row_ids")
    counts = torch.empty((2, d), dtype=torch.int64, device=dev)
    args = (_p(uvals64), _p(bg), K, d1, d2, _p(rs), rs.numel(), _p(cs), cs.numel(), bh, bw, _p(ranks), _p(col0), n_local, _p(row_ids))
    _call("pmd_export_csr", *args, 0, _p(counts[0]), _p(counts[1]), None, None, None, None, None, None, _stream())
    indptr = torch.zeros((2, d + 1), dtype=torch.int64, device=dev)
    for i in range(2):   # two 1-D scans (the device-wide scan; a (2, d) scan along dim 1 runs as two single-CTA scans: 0.5 ms)
        torch.cumsum(counts[i], dim=0, out=indptr[i, 1:])
    nnz_max = n_local * bh * bw + K * d
    cols = torch.empty((2, nnz_max), dtype=torch.int32, device=dev)
    vals64 = torch.empty(nnz_max, dtype=torch.float64, device=dev)
    vals32 = torch.empty(nnz_max, dtype=torch.float32, device=dev)
    _call("pmd_export_csr", *args, 1, None, None, _p(indptr[0]), _p(indptr[1]), _p(cols[0]), _p(vals64), _p(cols[1]), _p(vals32), _stream())
    return (indptr[0], cols[0], vals64), (indptr[1], cols[1], vals32)


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


# ---------------------------------------------------------------------------------------------
# K7 v3: host tables of pmd_project_stream
# ---------------------------------------------------------------------------------------------
PS_WARPS, PS_MAX_RW = 8, 48


def _split8(rk):
    """Split rk kept components into tasks of <= 8 components (sizes as equal as possible)."""
    n = (rk + 7) // 8
    base, extra = divmod(rk, n)
    return [base + (1 if i < extra else 0) for i in range(n)]


def _pack_slots(tasks, n_slots):
    """Greedy interval packing of tasks (sorted by first row) into ONE pass of n_slots task lists whose row ranges are
    disjoint and ascending.  Returns (slots, leftovers)."""
    slots = [[] for _ in range(n_slots)]
    left = []
    for tk in tasks:
        for s in slots:
            if not s or s[-1][0] + s[-1][2] <= tk[0]:
                s.append(tk)
                break
        else:
            left.append(tk)
    return slots, left


def _balance_slots(slots):
    """Permute warp slots so that the work (rows x width x padded comps) is spread evenly over the four SM
    sub-partitions (warp w issues on sub-partition w % 4)."""
    n = len(slots)
    load = [sum(tk[2] * tk[3] * tk[6] for tk in s) for s in slots]
    order = sorted(range(n), key=lambda i: -load[i])
    bucket_load, bucket_free = [0] * 4, [[b + 4 * j for j in range(n // 4)] for b in range(4)]
    out = [None] * n
    for i in order:
        b = min((b for b in range(4) if bucket_free[b]), key=lambda b: bucket_load[b])
        out[bucket_free[b].pop(0)] = slots[i]
        bucket_load[b] += load[i]
    return out


def _pack_passes(tasks, n_slots):
    """All passes needed for `tasks`: the first takes what fits; leftovers are clustered by contiguous row coverage so
    that every extra pass only streams the rows its tasks need.  Returns [(row0, row1, slots), ...]."""
    out = []
    slots, left = _pack_slots(tasks, n_slots)
    used = [tk for s in slots for tk in s]
    if used:
        out.append((min(tk[0] for tk in used), max(tk[0] + tk[2] for tk in used), _balance_slots(slots)))
    while left:
        cluster, end, rest = [], None, []
        for tk in left:
            if end is None or tk[0] < end:
                cluster.append(tk)
                end = tk[0] + tk[2] if end is None else max(end, tk[0] + tk[2])
            else:
                rest.append(tk)
        slots, more = _pack_slots(cluster, n_slots)
        used = [tk for s in slots for tk in s]
        out.append((min(tk[0] for tk in used), max(tk[0] + tk[2] for tk in used), slots))
        left = sorted(more + rest, key=lambda x: x[0])
    return out


def make_strips_py(row_starts, col_starts, bh, bw, d1, d2, ranks_host, col0_host, n_bg, G=None):
    """Pure-Python twin of make_strips (pmd_make_strips), kept as the executable specification for the tests.
    Host tables for pmd_project_stream (see include/pmd_sm100.h).  Blocks form the grid row_starts x col_starts
    (numbered row-major).  Returns a dict with int32 arrays items [n,8], slot_ptr, tasks [m,12], the order in which
    local tasks must be packed (`local8`, `local4`: arrays of (first column, n comps)), the background groups,
    `n_parts` and `max_rw`; or None when a single block column is already wider than the kernel's strip limit."""
    row_starts, col_starts = [int(x) for x in row_starts], [int(x) for x in col_starts]
    nbr, nbc = len(row_starts), len(col_starts)
    ranks = np.asarray(ranks_host, dtype=np.int64).reshape(nbr, nbc)
    col0 = np.asarray(col0_host, dtype=np.int64).reshape(nbr, nbc)
    if bw > PS_MAX_RW:
        return None
    bpix = bh * bw
    bg_groups = [(k0, min(8, n_bg - k0)) for k0 in range(0, n_bg, 8)]

    def build(g):
        local8, local4, items = [], [], []
        total_rw = 0
        for part, ca in enumerate(range(0, nbc, g)):
            cb = min(ca + g, nbc)
            c0, c1 = col_starts[ca], col_starts[cb - 1] + bw
            if c1 - c0 > PS_MAX_RW:
                return None
            core_end = col_starts[cb] if cb < nbc else d2
            tasks = []
            # background tasks first (they keep their slot for the whole strip); (by, bx, h, w, col, nc, ncp, kind, key)
            for gi, (k0, nc) in enumerate(bg_groups):
                tasks.append((0, 0, d1, core_end - c0, k0, nc, 8 if nc > 4 else 4, 1, gi))
            loc = []
            for a in range(nbr):
                for c in range(ca, cb):
                    first = int(col0[a, c])
                    for nc in _split8(int(ranks[a, c])):
                        loc.append((row_starts[a], col_starts[c] - c0, bh, bw, first, nc, 8 if nc > 4 else 4, 0, None))
                        first += nc
            loc.sort(key=lambda x: x[0])
            for (row0, row1, slots) in _pack_passes(tasks + loc, PS_WARPS):
                items.append(dict(c0=c0, rw=c1 - c0, slots=slots, bg_part=part, row0=row0, n_rows=row1 - row0))
                total_rw += ((c1 - c0 + 31) // 32 * 32) * (row1 - row0)
        return items, total_rw

    best = None
    for g in ([G] if G else range(1, 9)):
        res = build(g)
        if res is None:
            break
        if best is None or res[1] < best[1]:
            best = res
    if best is None:
        return None
    items, _ = best
    n8 = n4 = 0
    for itm in items:
        for s in itm["slots"]:
            for tk in s:
                if tk[7] == 0:
                    if tk[6] == 8:
                        n8 += 1
                    else:
                        n4 += 1
    base4 = n8 * bpix * 8
    base_bg = base4 + n4 * bpix * 4
    bg_off, off = [], base_bg
    for (k0, nc) in bg_groups:
        ncp = 8 if nc > 4 else 4
        bg_off.append(off)
        off += d1 * d2 * ncp
    local8, local4 = [], []
    it_arr, slot_ptr, task_arr = [], [], []
    for itm in items:
        it_arr.append((itm["c0"], itm["rw"], len(slot_ptr), itm["n_rows"], itm["bg_part"], itm["row0"], 0, 0))
        for s in itm["slots"]:
            slot_ptr.append(len(task_arr))
            for (by, bx, h, w, col, nc, ncp, kind, key) in s:
                if kind == 0:
                    if ncp == 8:
                        uoff = len(local8) * bpix * 8
                        local8.append((col, nc))
                    else:
                        uoff = base4 + len(local4) * bpix * 4
                        local4.append((col, nc))
                    urow = bw * ncp
                else:
                    uoff = bg_off[key] + itm["c0"] * ncp
                    urow = d2 * ncp
                task_arr.append((by, bx, h, w, col, nc, ncp, urow, uoff & 0xFFFFFFFF, uoff >> 32, kind, 0))
        slot_ptr.append(len(task_arr))
    u32 = lambda a, w: np.array(a, dtype=np.int64).astype(np.uint32).view(np.int32).reshape(-1, w)  # noqa: E731
    return dict(items=np.array(it_arr, dtype=np.int32).reshape(-1, 8), slot_ptr=np.array(slot_ptr, dtype=np.int32),
                tasks=u32(task_arr, 12), local8=np.array(local8, dtype=np.int64).reshape(-1, 2),
                local4=np.array(local4, dtype=np.int64).reshape(-1, 2), bg_groups=bg_groups, upack_floats=off,
                n_parts=1 + max(i["bg_part"] for i in items),
                max_rw=max(i["rw"] for i in items), n_items=len(items))


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def make_strips(row_starts, col_starts, bh, bw, d1, d2, ranks_host, col0_host, n_bg, G=None):
    """Host tables for pmd_project_stream, built by the native host routine pmd_make_strips (same result as
    make_strips_py).  Returns the same dict, or None when the geometry is not supported by the kernel."""
    rs = np.ascontiguousarray(row_starts, dtype=np.int32)
    cs = np.ascontiguousarray(col_starts, dtype=np.int32)
    ranks = np.ascontiguousarray(ranks_host, dtype=np.int64).reshape(-1)
    col0 = np.ascontiguousarray(col0_host, dtype=np.int64).reshape(-1)
    n_bg = int(n_bg)
    n_groups = (n_bg + 7) // 8
    cap_tasks = int(((ranks + 7) // 8).sum()) + len(cs) * n_groups + 8
    cap_items = len(cs) + cap_tasks
    items = np.zeros((cap_items, 8), np.int32)
    slot_ptr = np.zeros(cap_items * 9, np.int32)
    tasks = np.zeros((cap_tasks, 12), np.int32)
    local8 = np.zeros((cap_tasks, 2), np.int64)
    local4 = np.zeros((cap_tasks, 2), np.int64)
    counts = np.zeros(8, np.int64)
    rc = _lib.lib().pmd_make_strips(_np_ptr(rs), len(rs), _np_ptr(cs), len(cs), int(bh), int(bw), int(d1), int(d2), _np_ptr(ranks),
                                    _np_ptr(col0), n_bg, int(G or 0), _np_ptr(items), cap_items, _np_ptr(slot_ptr), _np_ptr(tasks),
                                    cap_tasks, _np_ptr(local8), _np_ptr(local4), _np_ptr(counts))
    _lib.check(rc, "pmd_make_strips")
    n_items, n_sp, n_tasks, n8, n4, n_parts, max_rw, upack_floats = (int(x) for x in counts)
    if n_items == 0:
        return None
    return dict(items=items[:n_items].copy(), slot_ptr=slot_ptr[:n_sp].copy(), tasks=tasks[:n_tasks].copy(),
                local8=local8[:n8].copy(), local4=local4[:n4].copy(),
                bg_groups=[(k0, min(8, n_bg - k0)) for k0 in range(0, n_bg, 8)], upack_floats=upack_floats, n_parts=n_parts,
                max_rw=max_rw, n_items=n_items)


def pack_strip_u(st, uvals32, bg, bpix):
    """upack for pmd_project_stream: per local task [pixel][padded comps], then per background group [pixel][comps]."""
    dev = uvals32.device
    n_local = uvals32.shape[0]
    upack = torch.zeros(st["upack_floats"], dtype=torch.float32, device=dev)
    uvp = torch.cat([uvals32, torch.zeros((1, bpix), dtype=torch.float32, device=dev)], dim=0)  # row n_local = zeros
    pos = 0
    for key, ncp in (("local8", 8), ("local4", 4)):
        lt = st[key]
        if len(lt):
            idx = lt[:, :1] + np.arange(ncp)[None, :]
            idx = np.where(np.arange(ncp)[None, :] < lt[:, 1:2], idx, n_local)
            g = uvp[torch.from_numpy(idx).to(dev)]  # (ntask, ncp, bpix)
            n = g.numel()
            upack[pos : pos + n] = g.permute(0, 2, 1).reshape(-1)
            pos += n
    K, d = bg.shape
    for (k0, nc) in st["bg_groups"]:
        ncp = 8 if nc > 4 else 4
        blk = torch.zeros((d, ncp), dtype=torch.float32, device=dev)
        blk[:, :nc] = bg[k0 : k0 + nc].t()
        upack[pos : pos + d * ncp] = blk.reshape(-1)
        pos += d * ncp
    assert pos == st["upack_floats"]
    return upack


def project_stream(movie2d, d2, st_dev, upack, mean, inv_std, z_local, z_bg):
    """One streaming pass: z_local[col, f] (local columns) and z_bg[k, f] (dense background columns) of
    U^T standardised movie.  st_dev: make_strips() tables with items/slot_ptr/tasks as device tensors."""
    t, d = movie2d.shape
    assert z_local.dtype == torch.float32 and (z_local.numel() == 0 or z_local.stride(1) == 1)
    K = z_bg.shape[0]
    parts = torch.empty((st_dev["n_parts"], K, t), dtype=torch.float32, device=movie2d.device)
    zl = z_local if z_local.numel() else parts
    _call("pmd_project_stream", _p(movie2d), movie_dtype_code(movie2d), t, d2, d, _p(st_dev["items"]), st_dev["n_items"],
          _p(st_dev["slot_ptr"]), _p(st_dev["tasks"]), st_dev["max_rw"], _p(upack), _p(mean), _p(inv_std), _p(zl),
          zl.stride(0) if z_local.numel() else t, _p(parts), t, K * t, _stream())
    z_bg[:, :t].copy_(parts.sum(dim=0))


# ---- K7 on the tensor cores (csrc/project_tc.cu, strips_tc_host.cu) ----------------------------------------------
TC_SLOTS, TC_CHUNK_BYTES = 32, 32768


def make_strips_tc(row_starts, col_starts, bh, bw, d1, d2, ranks_host, col0_host, n_bg, G=None):
    """Host tables for pmd_project_stream_tc (native routine pmd_make_strips_tc).  Returns a dict of numpy
    tables, or None when the geometry is not supported by the tensor-core kernel (callers then use
    make_strips / project_stream)."""
    rs = np.ascontiguousarray(row_starts, dtype=np.int32)
    cs = np.ascontiguousarray(col_starts, dtype=np.int32)
    ranks = np.ascontiguousarray(ranks_host, dtype=np.int64).reshape(-1)
    col0 = np.ascontiguousarray(col0_host, dtype=np.int64).reshape(-1)
    n_bg = int(n_bg)
    cap_tasks = int(((ranks + 3) // 4).sum()) + len(cs) * ((n_bg + 3) // 4) + 8
    cap_items = len(cs) + cap_tasks
    items = np.zeros((cap_items, 12), np.int32)
    slot_ptr = np.zeros(cap_items * (TC_SLOTS + 1), np.int32)
    tasks = np.zeros((cap_tasks, 8), np.int32)
    cap_events = cap_tasks + ((n_bg + 3) // 4) * (len(cs) + 4) * (int(d1) + 2)  # local tasks + background drains
    events = np.zeros((cap_events, 4), np.int32)
    counts = np.zeros(8, np.int64)
    rc = _lib.lib().pmd_make_strips_tc(_np_ptr(rs), len(rs), _np_ptr(cs), len(cs), int(bh), int(bw), int(d1), int(d2),
                                       _np_ptr(ranks), _np_ptr(col0), n_bg, int(G or 0), _np_ptr(items), cap_items,
                                       _np_ptr(slot_ptr), _np_ptr(tasks), cap_tasks, _np_ptr(events), cap_events, _np_ptr(counts))
    _lib.check(rc, "pmd_make_strips_tc")
    n_items, n_tasks, n_ev, chunks, n_parts, max_w8, g = (int(x) for x in counts[:7])
    if n_items == 0:
        return None
    items = items[:n_items].copy()
    item_of_row = np.concatenate([np.stack([np.full(int(it[3]), i, np.int32), np.arange(int(it[3]), dtype=np.int32)], axis=1)
                                  for i, it in enumerate(items)], axis=0)
    return dict(items=items, slot_ptr=slot_ptr[: n_items * (TC_SLOTS + 1)].copy(), tasks=tasks[:n_tasks].copy(),
                events=events[:n_ev].copy(), item_of_row=np.ascontiguousarray(item_of_row), chunks=chunks, n_parts=n_parts,
                max_w8=max_w8, G=g, n_items=n_items)


def pack_strips_tc(st_dev, uvals32, bg, bpix, d2):
    """Coefficient images (uint8 tensor, 32 KB per (item, row, 32-pixel chunk)) of pmd_project_stream_tc.
    st_dev: make_strips_tc() tables with items / slot_ptr / tasks / events / item_of_row as device tensors."""
    dev = st_dev["items"].device
    bimg = torch.empty(st_dev["chunks"] * TC_CHUNK_BYTES, dtype=torch.uint8, device=dev)
    uv = uvals32 if uvals32.numel() else torch.zeros((1, bpix), dtype=torch.float32, device=dev)
    d = bg.shape[1] if bg is not None and bg.numel() else d2
    _call("pmd_pack_strips_tc", _p(st_dev["items"]), _p(st_dev["item_of_row"]), st_dev["item_of_row"].shape[0],
          _p(st_dev["slot_ptr"]), _p(st_dev["tasks"]), _p(_req(uv, torch.float32, "uvals")), bpix,
          _p(bg) if bg is not None and bg.numel() else None, d, d2, _p(bimg), _stream())
    return bimg


def project_stream_tc_ok(movie2d, d2, mean, inv_std):
    """Alignment requirements of the tensor-core projection kernel."""
    ok = d2 % 4 == 0 and movie2d.data_ptr() % 16 == 0 and movie2d.stride(0) == movie2d.shape[1]
    for v in (mean, inv_std):
        ok = ok and (v is None or v.data_ptr() % 16 == 0)
    return ok


def project_stream_tc(movie2d, d2, st_dev, bimg, mean, inv_std, z_local, z_bg):
    """K7 on tcgen05: z_local[col, f] and z_bg[k, f] of U^T standardised movie in one streaming pass."""
    t, d = movie2d.shape
    assert z_local.dtype == torch.float32 and (z_local.numel() == 0 or z_local.stride(1) == 1)
    K = z_bg.shape[0]
    parts = torch.zeros((st_dev["n_parts"], max(K, 1), t), dtype=torch.float32, device=movie2d.device)
    zl = z_local if z_local.numel() else parts
    _call("pmd_project_stream_tc", _p(movie2d), movie_dtype_code(movie2d), t, d2, d, _p(st_dev["items"]), st_dev["n_items"],
          _p(st_dev["events"]), _p(bimg), _p(mean), _p(inv_std), _p(zl), zl.stride(0) if z_local.numel() else t, _p(parts), t,
          max(K, 1) * t, _stream())
    if K:
        z_bg[:, :t].copy_(parts[:, :K].sum(dim=0))


# ---- K7, movie operand in tensor memory (csrc/project_ts.cu, strips_ts_host.cu) ------------------------------------
TS_MAX_SLOTS = 48


def make_strips_ts(row_starts, col_starts, bh, bw, d1, d2, ranks_host, col0_host, n_bg, W=None, N=None):
    """Host tables for pmd_project_stream_ts (native routine pmd_make_strips_ts).  Returns a dict of numpy tables, or
    None when the geometry is not supported (callers then use the older kernels)."""
    rs = np.ascontiguousarray(row_starts, dtype=np.int32)
    cs = np.ascontiguousarray(col_starts, dtype=np.int32)
    ranks = np.ascontiguousarray(ranks_host, dtype=np.int64).reshape(-1)
    col0 = np.ascontiguousarray(col0_host, dtype=np.int64).reshape(-1)
    n_bg = int(n_bg)
    n_strips_max = (int(d2) + 31) // 32
    # a block gives a task to at most two strips
    cap_tasks = 2 * int(((ranks + 3) // 4).sum()) + n_strips_max * ((n_bg + 3) // 4) + 8
    cap_items = n_strips_max + cap_tasks
    items = np.zeros((cap_items, 12), np.int32)
    slot_ptr = np.zeros(cap_items * (TS_MAX_SLOTS + 1), np.int32)
    tasks = np.zeros((cap_tasks, 8), np.int32)
    cap_events = cap_tasks + ((n_bg + 3) // 4) * (n_strips_max + 4) * (int(d1) + 2)  # local tasks + background drains
    events = np.zeros((cap_events, 4), np.int32)
    counts = np.zeros(8, np.int64)
    rc = _lib.lib().pmd_make_strips_ts(_np_ptr(rs), len(rs), _np_ptr(cs), len(cs), int(bh), int(bw), int(d1), int(d2),
                                       _np_ptr(ranks), _np_ptr(col0), n_bg, int(W or 0), int(N or 0), _np_ptr(items), cap_items,
                                       _np_ptr(slot_ptr), _np_ptr(tasks), cap_tasks, _np_ptr(events), cap_events, _np_ptr(counts))
    _lib.check(rc, "pmd_make_strips_ts")
    n_items, n_tasks, n_ev, chunks, n_parts, w, n, tiles = (int(x) for x in counts[:8])
    if n_items == 0:
        return None
    items = items[:n_items].copy()
    item_of_row = np.concatenate([np.stack([np.full(int(it[3]), i, np.int32), np.arange(int(it[3]), dtype=np.int32)], axis=1)
                                  for i, it in enumerate(items)], axis=0)
    tk = tasks[:n_tasks]
    return dict(items=items, slot_ptr=slot_ptr[: n_items * (TS_MAX_SLOTS + 1)].copy(), tasks=tk.copy(),
                events=events[:n_ev].copy(), item_of_row=np.ascontiguousarray(item_of_row), chunks=chunks, n_parts=n_parts,
                W=w, N=n, tiles=tiles, n_items=n_items, has_shared=bool((tk[:, 6] == 2).any()))


def pack_strips_ts(st_dev, uvals32, bg, inv_std, bpix, d2):
    """Coefficient images (uint8 tensor, 2 N 128 bytes per (item, row, 32-pixel chunk)) of pmd_project_stream_ts, with
    1 / std folded in.  st_dev: make_strips_ts() tables with the arrays as device tensors."""
    dev = st_dev["items"].device
    n = st_dev["N"]
    bimg = torch.empty(st_dev["chunks"] * 2 * n * 128, dtype=torch.uint8, device=dev)
    uv = uvals32 if uvals32.numel() else torch.zeros((1, bpix), dtype=torch.float32, device=dev)
    have_bg = bg is not None and bg.numel() > 0
    d = bg.shape[1] if have_bg else (inv_std.numel() if inv_std is not None else d2)
    if inv_std is not None:
        _req(inv_std, torch.float32, "inv_std")
    _call("pmd_pack_strips_ts", _p(st_dev["items"]), _p(st_dev["item_of_row"]), st_dev["item_of_row"].shape[0],
          _p(st_dev["slot_ptr"]), _p(st_dev["tasks"]), _p(_req(uv, torch.float32, "uvals")), bpix, _p(bg) if have_bg else None,
          _p(inv_std), d, d2, n, _p(bimg), _stream())
    return bimg


def project_stream_ts_ok(movie2d, d2, mean):
    """Alignment requirements of the TMA-fed projection kernel."""
    pitch = movie2d.shape[1] * movie2d.element_size()
    # every TMA box starts at a multiple of 32 pixels of an image row: rows and frames must start on 16-byte boundaries
    ok = (d2 % 4 == 0 and movie2d.data_ptr() % 16 == 0 and movie2d.stride(0) == movie2d.shape[1] and pitch % 16 == 0
          and (d2 * movie2d.element_size()) % 16 == 0)
    return ok and (mean is None or mean.data_ptr() % 16 == 0)


def project_stream_ts(movie2d, d2, st_dev, bimg, mean, z_local, z_bg, mark=None):
    """K7 (TMA + tensor-memory operand): z_local[col, f] and z_bg[k, f] of U^T standardised movie in one streaming pass.
    `bimg` must have been packed with the inv_std this projection is meant to apply (pack_strips_ts).
    mark: optional callable(name) for CUDA-event marks: "projection.prep" is recorded right before the kernel launch (table
    uploads, coefficient images and the zero fills lie before it), "projection.stream" right after it -- the kernel alone."""
    t, d = movie2d.shape
    assert z_local.dtype == torch.float32 and (z_local.numel() == 0 or z_local.stride(1) == 1)
    K = z_bg.shape[0]
    parts = torch.zeros((st_dev["n_parts"], max(K, 1), t), dtype=torch.float32, device=movie2d.device)
    zl = z_local if z_local.numel() else parts
    if st_dev["has_shared"] and z_local.numel():
        z_local[:, :t].zero_()   # blocks shared by two strips are accumulated by two atomic adds
    if mark is not None:
        mark("projection.prep")
    _call("pmd_project_stream_ts", _p(movie2d), movie_dtype_code(movie2d), t, d2, d, _p(st_dev["items"]), st_dev["n_items"],
          _p(st_dev["events"]), _p(bimg), st_dev["N"], _p(mean), _p(zl), zl.stride(0) if z_local.numel() else t, _p(parts), t,
          max(K, 1) * t, _stream())
    if mark is not None:
        mark("projection.stream")
    if K:
        z_bg[:, :t].copy_(parts[:, :K].sum(dim=0))
