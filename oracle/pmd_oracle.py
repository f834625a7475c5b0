"""CPU oracle for the localmd PMD hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (localmd_b200) never imports it and has no CPU fallback.

It restates, in NumPy/SciPy float32, the algorithm of apasarkar/localmd v0.0.4 (reference checkout
at /root/reference, all citations `file:line` relative to it).  The arithmetic of the reference lives
in JAX (unpinned, absent from this image); every jnp call site is restated with its NumPy/LAPACK
equivalent (jnp.linalg.qr -> numpy.linalg.qr reduced, jnp.linalg.svd -> numpy.linalg.svd,
svd(hermitian=True) -> numpy's eigh-based hermitian svd, jax.scipy.signal.welch -> scipy.signal.welch).

PINNING: tests/test_oracle_golden.py checks this module against fixtures produced by executing the
UNMODIFIED reference source over a NumPy-backed JAX stand-in (tests/golden/make_golden.py,
tests/golden/jax_shim/) on the same inputs and the same random draws.  That pins the restated
control flow, reshape orders, index conventions and sparse assembly to the reference's own code;
XLA-CPU float32 numerics themselves remain unpinned (the reference's tests hold no numeric
assertions and no golden vectors).

Randomness is never drawn here implicitly: every random quantity is an explicit argument
(`Draws`), so the CUDA path and the oracle can be fed identical numbers.
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import scipy.signal
import scipy.sparse as sp

F32 = np.float32  # working precision of the restatement; see precision()


@contextlib.contextmanager
def precision(dtype):
    """Run the restated algorithm in another working precision.  precision(np.float64) evaluates the
    reference's algorithm (same control flow, same draws) in double: the tests use it as the
    exact-arithmetic answer, to separate 'differs from the reference algorithm' from 'differs by the
    reference's own float32 rounding'."""
    global F32
    old = F32
    F32 = np.dtype(dtype).type
    try:
        yield
    finally:
        F32 = old


# --------------------------------------------------------------------------------------------
# random draws (explicit)
# --------------------------------------------------------------------------------------------
def normal_from_seed(seed: int, shape) -> np.ndarray:
    """The documented Gaussian generator shared by the golden generator's jax.random.normal stand-in,
    the oracle and the GPU parity tests: Generator(PCG64(seed mod 2**32)).standard_normal(float32)."""
    rng = np.random.Generator(np.random.PCG64(int(seed) & 0xFFFFFFFF))
    return rng.standard_normal(tuple(int(s) for s in shape), dtype=np.float32)


@dataclass
class Draws:
    """Every random quantity localmd_decomposition consumes, in the reference's consumption order.

    bg_frames      frame ids for the background rSVD       (pmd_loader.py:303-306, np.random.choice)
    bg_sketch      (n_bg, background_rank+10) Gaussian      (pmd_loader.py:56)
    init_frames    frame ids used for the block fits        (decomposition.py:681-693)
    sim_noise      list of (bh,bw,t) Gaussians, 250 long    (decomposition.py:127)   } or `thresholds`
    sim_sketch     list of (t, 11) Gaussians, 250 long      (decomposition.py:62)    }
    thresholds     optional (spatial, temporal) overriding the simulation
    block_sketches list over blocks (dim1 outer, dim2 inner) of lists over windows of (t', r+10) Gaussians
                                                            (decomposition.py:475, 62)
    prune_sketch   (t, int(min(R,t)*factor)) Gaussian        (decomposition.py:870-872); a callable
                   shape->array is accepted because R is data dependent
    """

    bg_frames: Optional[Sequence[int]] = None
    bg_sketch: Optional[np.ndarray] = None
    init_frames: Optional[Sequence[int]] = None
    sim_noise: Optional[Sequence[np.ndarray]] = None
    sim_sketch: Optional[Sequence[np.ndarray]] = None
    thresholds: Optional[Tuple[float, float]] = None
    block_sketches: Optional[Sequence] = None
    prune_sketch: Optional[object] = None


# --------------------------------------------------------------------------------------------
# a1: mean + noise normaliser            pmd_loader.py:203-291, preprocessing_utils.py:10-40
# --------------------------------------------------------------------------------------------
def welch_noise_estimate(traces: np.ndarray) -> np.ndarray:
    """preprocessing_utils.py:28-37 for a batch of traces (npix, n), n >= 256, float32.
    welch(trace, noverlap=128) with defaults fs=1, periodic Hann, nperseg=256, constant detrend,
    one-sided density, mean over segments; then bins 65..128, x0.5, mean over the 64 bins, sqrt."""
    traces = np.asarray(traces, dtype=F32)
    _, pxx = scipy.signal.welch(traces, noverlap=128, axis=-1)
    pxx = pxx.astype(F32)
    start = int(256 / 4 + 1)
    end = int(256 / 2 + 1)
    vals = pxx[..., start:end] * F32(0.5)
    return np.sqrt(np.sum(vals, axis=-1, dtype=F32) / F32(end - start)).astype(F32)


def welch_noise_estimate_explicit(traces: np.ndarray) -> np.ndarray:
    """Same quantity written out from the definition (float64), used by the tests to show that the
    band power only needs the DFT bins 65..128 of each Hann-windowed, mean-detrended 256-sample
    segment at hop 128 -- the form the CUDA stats kernel computes."""
    x = np.asarray(traces, dtype=np.float64)
    n = x.shape[-1]
    nseg = (n - 128) // 128
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(256) / 256.0)
    scale = 1.0 / np.sum(w * w)  # fs = 1
    acc = np.zeros(x.shape[:-1] + (129,))
    for s in range(nseg):
        seg = x[..., 128 * s : 128 * s + 256]
        seg = seg - seg.mean(axis=-1, keepdims=True)
        spec = np.fft.rfft(seg * w, axis=-1)
        p = (spec.real**2 + spec.imag**2) * scale
        p[..., 1:128] *= 2.0  # one-sided: DC and Nyquist are not doubled
        acc += p
    acc /= nseg
    vals = acc[..., 65:129] * 0.5
    return np.sqrt(vals.sum(axis=-1) / 64.0)


def mean_and_noise(movie, compute_normalizer: bool = True, frame_constant: int = 1024, min_allowed_frames: int = 256):
    """pmd_loader.py:203-291.  `movie` is anything with .shape (T,d1,d2) and movie[list] -> (n,d1,d2).
    Returns (mean_img, std_img) float32 (d1,d2).  The reference's pixel tiling (228-243, 260-280) is a
    memory device: every per-pixel result is independent of it, so it is not restated."""
    T, d1, d2 = movie.shape
    flag = bool(compute_normalizer) and T >= min_allowed_frames
    overall_mean = np.zeros((d1, d2), dtype=F32)
    overall_norm = np.zeros((d1, d2), dtype=F32) if flag else np.ones((d1, d2), dtype=F32)
    starts = list(range(0, T, frame_constant))
    n_var = 0
    for s in starts:
        e = min(s + frame_constant, T)
        chunk = np.asarray(movie[list(range(s, e))]).astype(F32)  # (n, d1, d2)
        n = chunk.shape[0]
        if n >= min_allowed_frames:
            n_var += 1
        # get_mean_and_noise / get_mean_chunk: float32 sum over the chunk divided by the TOTAL frame count
        mean_chunk = (np.sum(chunk, axis=0, dtype=F32) / F32(T)).astype(F32)
        overall_mean += mean_chunk.astype(np.float64)
        if flag and n >= min_allowed_frames:
            traces = chunk.reshape(n, d1 * d2).T  # (npix, n), C-order pixel id (per-pixel op)
            noise = welch_noise_estimate(traces).reshape(d1, d2)
            overall_norm += noise.astype(np.float64) / len(starts)
    if flag and n_var != 0:
        overall_norm *= len(starts) / n_var
        overall_norm[overall_norm == 0] = 1
    return overall_mean, overall_norm


# --------------------------------------------------------------------------------------------
# randomized SVDs
# --------------------------------------------------------------------------------------------
def _rsvd_core(a: np.ndarray, omega: np.ndarray):
    """Shared body of decomposition.py:62-67 and pmd_loader.py:56-62 (float32)."""
    a = np.asarray(a, dtype=F32)
    omega = np.asarray(omega, dtype=F32)
    projected = a @ omega
    q, _ = np.linalg.qr(projected)
    b = q.T @ a
    u, s, v = np.linalg.svd(b, full_matrices=False)
    return (q @ u).astype(F32), s.astype(F32), v.astype(F32)


def _dynamic_slice_cols(m: np.ndarray, n: int) -> np.ndarray:
    if n > m.shape[1]:
        raise TypeError("rank larger than available columns (jax.lax.dynamic_slice would fail)")
    return m[:, :n]


def truncated_random_svd_block(a, omega, rank):
    """decomposition.py:37-73: returns (u[:, :rank], s[:rank], v[:rank])."""
    u, s, v = _rsvd_core(a, omega)
    if rank > v.shape[0] or rank > u.shape[1]:
        raise TypeError("rank larger than the sketch allows (jax.lax.dynamic_slice would fail)")
    return u[:, :rank], s[:rank], v[:rank, :]


def background_basis(movie, mean_img, std_img, bg_frames, bg_sketch, background_rank, order="F"):
    """pmd_loader.py:293-314 (+46-68).  Returns the (d, background_rank) float32 orthonormal spatial
    background basis, rows in `order` pixel numbering; (d,1) zeros when background_rank <= 0."""
    T, d1, d2 = movie.shape
    if background_rank <= 0:
        return np.zeros((d1 * d2, 1), dtype=F32)
    crop = np.asarray(movie[list(bg_frames)]).astype(F32).transpose(1, 2, 0)
    crop = crop - mean_img[:, :, None]
    crop = crop / std_img[:, :, None]
    crop = crop.astype(F32)
    a = crop.reshape((-1, crop.shape[-1]), order=order)
    u, s, v = _rsvd_core(a, bg_sketch)
    return _dynamic_slice_cols(u, background_rank).astype(F32)


# --------------------------------------------------------------------------------------------
# a8: roughness statistics + rank rule            evaluation.py:84-222
# --------------------------------------------------------------------------------------------
def spatial_roughness_stat(u: np.ndarray) -> np.float32:
    """evaluation.py:84-111 for one (d1,d2) image, float32."""
    u = np.ascontiguousarray(u, dtype=F32)
    vert = np.abs(u[1:, :] - u[:-1, :])
    horiz = np.abs(u[:, :-1] - u[:, 1:])
    avg_diff = (np.sum(vert, dtype=F32) + np.sum(horiz, dtype=F32)) / F32(vert.size + horiz.size)
    with np.errstate(divide="ignore", invalid="ignore"):
        return F32(avg_diff / np.mean(np.abs(u), dtype=F32))


def temporal_roughness_stat(v: np.ndarray) -> np.float32:
    """evaluation.py:114-126 for one trace, float32."""
    v = np.ascontiguousarray(v, dtype=F32)
    num = np.mean(np.abs(v[:-2] + v[2:] - F32(2) * v[1:-1]), dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return F32(num / np.mean(np.abs(v), dtype=F32))


def fitness_decisions(u3: np.ndarray, v: np.ndarray, thr_s, thr_t):
    """evaluation.py:133-192: u3 (d1,d2,r), v (r,t).  Thresholds are cast to float32 as jit does.
    Returns (good int32 (r,), spatial stats (r,), temporal stats (r,))."""
    r = u3.shape[2]
    ss = np.array([spatial_roughness_stat(u3[:, :, c]) for c in range(r)], dtype=F32)
    ts = np.array([temporal_roughness_stat(v[c]) for c in range(r)], dtype=F32)
    good = ((ss < F32(thr_s)) & (ts < F32(thr_t))).astype(np.int32)
    return good, ss, ts


def filter_by_failures(decisions: np.ndarray, max_consecutive_failures: int) -> np.ndarray:
    """evaluation.py:195-222: a failing component is kept until the running count of consecutive
    failures reaches the limit; everything after that is dropped."""
    d = np.array(decisions, dtype=bool).copy()
    fails = 0
    dead = False
    for k in range(d.shape[0]):
        if dead:
            d[k] = False
        elif not d[k]:
            fails += 1
            d[k] = True
            if fails == max_consecutive_failures:
                dead = True
        else:
            fails = 0
    return d


# --------------------------------------------------------------------------------------------
# a5: threshold simulation            decomposition.py:76-189
# --------------------------------------------------------------------------------------------
def rank_simulation(noise: np.ndarray, sketch: np.ndarray, num_comps: int = 1):
    """decomposition.py:102-131 + 76-99: rank-`num_comps` rSVD of a pure-noise block at full
    resolution, then the two roughness statistics of (u, s*v)."""
    d1, d2, t = noise.shape
    block_2d = np.reshape(np.asarray(noise, dtype=F32), (d1 * d2, t), order="F")
    u, s, v = truncated_random_svd_block(block_2d, sketch, num_comps)
    v = s[:, None] * v
    u3 = np.reshape(u, (d1, d2, u.shape[1]), order="F")
    ss = np.array([spatial_roughness_stat(u3[:, :, c]) for c in range(u3.shape[2])], dtype=F32)
    ts = np.array([temporal_roughness_stat(v[c]) for c in range(v.shape[0])], dtype=F32)
    return ss, ts


def threshold_heuristic(sim_noise, sim_sketch, percentile_threshold=5, num_comps=1):
    """decomposition.py:147-189."""
    sl, tl = [], []
    for noise, sk in zip(sim_noise, sim_sketch):
        x, y = rank_simulation(noise, sk, num_comps)
        sl.append(x)
        tl.append(y)
    thr_s = np.percentile(np.array(sl).flatten(), percentile_threshold)
    thr_t = np.percentile(np.array(tl).flatten(), percentile_threshold)
    return thr_s, thr_t


# --------------------------------------------------------------------------------------------
# a3: init-frame selection            decomposition.py:528-569, 678-693
# --------------------------------------------------------------------------------------------
def window_chunk_candidates(frame_range: int, total_frames: int, window_chunks: int) -> Tuple[np.ndarray, int]:
    """decomposition.py:546-555: candidate window starts and how many the caller must choose."""
    if frame_range > total_frames:
        raise ValueError("Requested more frames than available")
    if window_chunks > frame_range:
        raise ValueError("The size of each temporal chunk is bigger than frame range")
    num_intervals = math.ceil(frame_range / window_chunks)
    avail = np.arange(0, total_frames, window_chunks)
    if avail[-1] > total_frames - window_chunks:
        avail[-1] = total_frames - window_chunks
    return avail, num_intervals


def frames_from_starts(starting_points, total_frames: int, window_chunks: int) -> List[int]:
    """decomposition.py:559-569."""
    out: List[int] = []
    for k in np.sort(np.asarray(starting_points)):
        out.extend(range(int(k), int(min(k + window_chunks, total_frames))))
    return out


# --------------------------------------------------------------------------------------------
# a4: standardise + background removal on the init frames      pmd_loader.py:348-389
# --------------------------------------------------------------------------------------------
def temporal_crop_with_filter(movie, frames, mean_img, std_img, spatial_basis, order, batch_size):
    T, d1, d2 = movie.shape
    crop = np.asarray(movie[list(frames)]).astype(F32).transpose(1, 2, 0)
    basis_r = spatial_basis.reshape((d1, d2, -1), order=order)
    out = np.zeros(crop.shape)  # float64 container, float32 values (pmd_loader.py:354-357)
    temporal_basis = np.zeros((basis_r.shape[2], crop.shape[2]))
    basis_2d = np.reshape(basis_r, (d1 * d2, basis_r.shape[2]), order="F").astype(F32)
    start = 0
    for _ in range(math.ceil(crop.shape[2] / batch_size)):
        end = min(crop.shape[2], start + batch_size)
        x = crop[:, :, start:end] - mean_img[:, :, None]
        x = (x / std_img[:, :, None]).astype(F32)
        x2 = np.reshape(x, (d1 * d2, x.shape[2]), order="F")
        tp = basis_2d.T @ x2
        x2 = x2 - basis_2d @ tp
        out[:, :, start:end] = np.reshape(x2, (d1, d2, -1), order="F")
        temporal_basis[:, start:end] = tp
        start += batch_size
    return out, temporal_basis


# --------------------------------------------------------------------------------------------
# a6: tiling + pyramid weights            decomposition.py:698, 723-754
# --------------------------------------------------------------------------------------------
def update_block_sizes(blocks, fov_shape, min_block_value=10):
    """decomposition.py:572-613."""
    if blocks[0] < min_block_value or blocks[1] < min_block_value:
        raise ValueError(
            "One of the block dimensions was less than min allowed value of {}, "
            "set to a larger value".format(min_block_value)
        )
    return [min(blocks[0], fov_shape[0]), min(blocks[1], fov_shape[1])]


def check_fov_size(fov_dims, min_allowed_value=10):
    """decomposition.py:616-635."""
    for k in fov_dims:
        if k < min_allowed_value:
            raise ValueError(
                "At least one FOV dimension is lower than {}, too small to process".format(min_allowed_value)
            )


def tile_starts(n: int, b: int) -> List[int]:
    """decomposition.py:698, 723-739: stride b - ceil(b/2), plus a forced last start n-b."""
    overlap = math.ceil(b / 2)
    it = list(range(0, n - b + 1, b - overlap))
    if it[-1] != n - b and n - b != 0:
        it.append(n - b)
    return it


def pyramid_weights(bh: int, bw: int) -> np.ndarray:
    """decomposition.py:742-750 (only well defined for even bh, bw -- odd sizes raise in the reference)."""
    w = np.ones((bh, bw), dtype=F32)
    hbh, hbw = bh // 2, bw // 2
    w[:hbh, :hbw] += np.minimum(np.tile(np.arange(0, hbw), (hbh, 1)), np.tile(np.arange(0, hbh), (hbw, 1)).T)
    w[:hbh, hbw:] = np.fliplr(w[:hbh, :hbw])
    w[hbh:, :] = np.flipud(w[:hbh, :])
    return w


# --------------------------------------------------------------------------------------------
# a7: per-block decomposition            decomposition.py:192-330, 333-407, 410-525
# --------------------------------------------------------------------------------------------
def downsample_average_pooling(arr: np.ndarray, n: int) -> np.ndarray:
    """decomposition.py:192-232: n x n sum pooling with XLA 'SAME' padding divided by the in-bounds count."""
    arr = np.asarray(arr, dtype=F32)
    d1, d2, t = arr.shape

    def pads(m):
        out = -(-m // n)
        total = max((out - 1) * n + n - m, 0)
        return total // 2, total - total // 2, out

    l1, h1, o1 = pads(d1)
    l2, h2, o2 = pads(d2)
    xp = np.pad(arr, ((l1, h1), (l2, h2), (0, 0)))
    cp = np.pad(np.ones((d1, d2, 1), dtype=F32), ((l1, h1), (l2, h2), (0, 0)))
    acc = np.zeros((o1, o2, t), dtype=F32)
    cnt = np.zeros((o1, o2, 1), dtype=F32)
    for a in range(n):
        for b in range(n):
            acc = acc + xp[a : a + (o1 - 1) * n + 1 : n, b : b + (o2 - 1) * n + 1 : n, :]
            cnt = cnt + cp[a : a + (o1 - 1) * n + 1 : n, b : b + (o2 - 1) * n + 1 : n, :]
    return (acc / cnt).astype(F32)


def single_block_md(block, sketch, max_rank, taf, saf, thr_s, thr_t, spatial_denoiser=None, temporal_denoiser=None):
    """decomposition.py:235-330.  block (d1,d2,t) float32, sketch (t//taf, max_rank+10).
    Returns (u (d1,d2,r), good (r,), v (r,t), (spatial stats, temporal stats))."""
    block = np.asarray(block, dtype=F32)
    d1, d2, t = block.shape
    ds = downsample_average_pooling(block, saf)
    d1n, d2n = ds.shape[0], ds.shape[1]
    ds_ta = np.mean(np.reshape(ds, (d1n * d2n, taf, t // taf), order="F"), axis=1, dtype=F32)
    u_ds = truncated_random_svd_block(ds_ta, sketch, max_rank)[0]
    v_ds = u_ds.T @ np.reshape(ds, (d1n * d2n, t), order="F")
    if temporal_denoiser is not None:
        v_ds = np.asarray(temporal_denoiser(v_ds), dtype=F32)
    v_basis = np.linalg.svd(v_ds, full_matrices=False)[2]
    block_2d = np.reshape(block, (d1 * d2, t), order="F")
    s_proj = block_2d @ v_basis.T
    if spatial_denoiser is not None:
        tmp = np.reshape(s_proj, (d1, d2, v_basis.shape[0]), order="F").transpose(2, 0, 1)
        tmp = np.asarray(spatial_denoiser(tmp), dtype=F32)
        s_proj = tmp.transpose(1, 2, 0).reshape((d1 * d2, v_basis.shape[0]), order="F")
    u_final = np.linalg.svd(s_proj, full_matrices=False)[0]
    v_new = u_final.T @ block_2d
    v_left, v_sing, v_right = np.linalg.svd(v_new, full_matrices=False)
    u_final = u_final @ v_left
    v_final = (v_sing[:, None] * v_right).astype(F32)
    u3 = np.reshape(u_final, (d1, d2, u_final.shape[1]), order="F").astype(F32)
    good, ss, ts = fitness_decisions(u3, v_final, thr_s, thr_t)
    return u3, good, v_final, (ss, ts)


def single_residual_block_md(block, existing, sketch, max_rank, taf, thr_s, thr_t):
    """decomposition.py:333-387 (window_chunks < frame_range only)."""
    block = np.asarray(block, dtype=F32)
    d1, d2, t = block.shape
    block_2d = np.reshape(block, (d1 * d2, t), order="F")
    ex = np.reshape(np.asarray(existing, dtype=F32), (d1 * d2, existing.shape[2]), order="F")
    block_2d = block_2d - ex @ (ex.T @ block_2d)
    avg = np.mean(np.reshape(block_2d, (d1 * d2, taf, t // taf), order="F"), axis=1, dtype=F32)
    u = truncated_random_svd_block(avg, sketch, max_rank)[0]
    v = u.T @ block_2d
    u3 = np.reshape(u, (d1, d2, u.shape[1]), order="F").astype(F32)
    good, ss, ts = fitness_decisions(u3, v, thr_s, thr_t)
    return u3, good, v, (ss, ts)


def windowed_pmd(
    window_length, block, max_rank, thr_s, thr_t, mcf, taf, saf, sketches, spatial_denoiser=None, temporal_denoiser=None
):
    """decomposition.py:410-525.  `sketches` holds one Gaussian per visited window.
    Returns (spatial (d1,d2,rank) float64, temporal (rank,t) float32, per-window diagnostics)."""
    block = np.asarray(block, dtype=F32)
    d1, d2, window_range = block.shape
    if window_length > window_range:
        window_length = window_range
    start_points = list(range(0, window_range, window_length))
    if len(start_points) > 0 and start_points[-1] + window_length > window_range:
        start_points[-1] = window_range - window_length
    final_spatial = np.zeros((d1, d2, max_rank))
    remaining = max_rank
    counter = 0
    diags = []
    for wi, k in enumerate(start_points):
        subset = block[:, :, k : k + window_length]
        if k == 0 or counter == 0:
            comps, dec, _, stats = single_block_md(
                subset, sketches[wi], max_rank, taf, saf, thr_s, thr_t, spatial_denoiser, temporal_denoiser
            )
        else:
            comps, dec, _, stats = single_residual_block_md(
                subset, final_spatial, sketches[wi], max_rank, taf, thr_s, thr_t
            )
        keep = filter_by_failures(dec.flatten() > 0, mcf)
        cropped = comps[:, :, keep]
        cropped = cropped[:, :, : min(cropped.shape[2], remaining)]
        final_spatial[:, :, counter : counter + cropped.shape[2]] = cropped
        counter += cropped.shape[2]
        diags.append(dict(good=dec, stats=stats, kept=cropped.shape[2]))
        if counter == max_rank:
            break
        remaining = max_rank - counter
    # get_temporal_projector (390-407) on the zero-padded (d1,d2,max_rank) basis
    proj = np.reshape(final_spatial.astype(F32), (d1 * d2, max_rank), order="F").T @ np.reshape(
        block, (d1 * d2, window_range), order="F"
    )
    return final_spatial[:, :, :counter], proj[:counter, :].astype(F32), diags


# --------------------------------------------------------------------------------------------
# a12 / a14: Gram-based whitening and final SVD            decomposition.py:936-1137
# --------------------------------------------------------------------------------------------
def _hermitian_svd(g: np.ndarray):
    u, s, _ = np.linalg.svd(np.asarray(g, dtype=F32), full_matrices=False, hermitian=True)
    return u.astype(F32), s.astype(F32)


def fewer_rows_svd_routine(data):
    """decomposition.py:1063-1099."""
    data = np.asarray(data, dtype=F32)
    left, vals = _hermitian_svd(data @ data.T)
    sing = np.sqrt(vals)
    div = np.where(sing == 0, F32(1), sing)
    right = (left.T @ data) / div[:, None]
    return left, sing, right.astype(F32)


def fewer_columns_svd_routine(data):
    """decomposition.py:1102-1137."""
    data = np.asarray(data, dtype=F32)
    right_t, vals = _hermitian_svd(data.T @ data)
    sing = np.sqrt(vals)
    div = np.where(sing == 0, F32(1), sing)
    left = data @ (right_t / div[None, :])
    return left.astype(F32), sing, right_t.T


def projected_svd(projection, data):
    """decomposition.py:1013-1060."""
    d1, d2 = data.shape
    if d1 <= d2:
        left, sing, right = fewer_rows_svd_routine(data)
    else:
        left, sing, right = fewer_columns_svd_routine(data)
    return (np.asarray(projection, dtype=F32) @ left).astype(F32), sing, right


def compute_lowrank_factorized_svd(u, v, only_left=False):
    """decomposition.py:936-1010.  u sparse (d,R) float64, v dense (R,t')."""
    u = sp.csr_matrix(u)
    ut_u = u.T.dot(u)
    right_mat = v if u.shape[1] > v.shape[1] else np.eye(u.shape[1])
    ut_ur = ut_u.dot(right_mat)
    g = np.asarray(right_mat, dtype=F32).T @ np.asarray(ut_ur, dtype=F32)
    vecs, vals = _hermitian_svd(g)
    good = vals > 0
    vecs, vals = vecs[:, good], vals[good]
    mix = (np.asarray(right_mat, dtype=F32) @ vecs).astype(F32)
    mix /= np.sqrt(vals)[None, :]
    if only_left:
        return mix
    new_temporal = mix.T @ np.asarray(ut_u.dot(v), dtype=F32)
    return projected_svd(mix, new_temporal)


# --------------------------------------------------------------------------------------------
# a13: full-movie projection            pmd_loader.py:71-108, 316-346, 392-414
# --------------------------------------------------------------------------------------------
def frame_batches(T: int, batch_size: int) -> List[Tuple[int, int]]:
    """FrameDataloader (pmd_loader.py:71-108): the last item absorbs the remainder."""
    chunks = math.ceil(T / batch_size)
    n = max(1, chunks - 1)
    out = []
    for index in range(n):
        start = index * batch_size
        end = T if index == max(0, chunks - 2) else start + batch_size
        out.append((start, end))
    return out


def v_projection(movie, u, mix, mean_img, std_img, order, batch_size):
    T, d1, d2 = movie.shape
    sparse = sp.csr_matrix(u.T).astype(F32)
    dense = np.asarray(mix, dtype=F32).T
    mean_r = mean_img.reshape((-1, 1), order=order)
    std_r = std_img.reshape((-1, 1), order=order)
    outs = []
    for s, e in frame_batches(T, batch_size):
        data = np.asarray(movie[list(range(s, e))]).astype(F32).transpose(1, 2, 0)
        data = np.reshape(data, (-1, data.shape[2]), order=order)
        centered = ((data - mean_r) / std_r).astype(F32)
        outs.append(dense @ np.asarray(sparse @ centered, dtype=F32))
    return np.concatenate(outs, axis=1).astype(F32)


# --------------------------------------------------------------------------------------------
# a15: PMDArray            pmdarray.py:7-171
# --------------------------------------------------------------------------------------------
class PMDArrayOracle:
    """pmdarray.py:7-171 restated.  (The 2-key form is broken in the reference, pmdarray.py:146-148;
    here it behaves as evidently intended: arr[f, rows] == arr[f, rows, :].)"""

    def __init__(self, u, r, s, v, data_shape, data_order, mean_img, std_img):
        self.order = data_order
        self.num_frames, self.fov_dim1, self.fov_dim2 = data_shape
        self.u = sp.csr_matrix(u)
        self.r, self.s, self.v = r, s, v
        self._combined = (self.r * self.s[None, :]).dot(self.v)
        self.mean_img, self.var_img = mean_img, std_img
        self.row_indices = np.arange(self.fov_dim1 * self.fov_dim2).reshape(
            (self.fov_dim1, self.fov_dim2), order=self.order
        )

    @property
    def shape(self):
        return (self.num_frames, self.fov_dim1, self.fov_dim2)

    @staticmethod
    def _l(e):
        return [e] if isinstance(e, (int, np.integer)) else e

    def __getitem__(self, key):
        if key is None:
            raise ValueError("Cannot use None for indexing")
        if not isinstance(key, tuple):
            key = (key,)
        if len(key) > 3:
            raise ValueError("Too many values to unpack in __getitem__")
        key = tuple(key) + (slice(None),) * (3 - len(key))
        if any(k is None for k in key):
            raise ValueError("Cannot use None for indexing")
        k1, k2 = self._l(key[1]), self._l(key[2])
        rows = self.row_indices[k1, k2]
        mean_used, var_used = self.mean_img[k1, k2], self.var_img[k1, k2]
        spatial = self.u[rows.reshape((-1,), order=self.order)]
        temporal = self._combined[:, self._l(key[0])]
        out = spatial.dot(temporal)
        out = out.reshape(rows.shape + (-1,), order=self.order) * np.expand_dims(
            var_used, axis=var_used.ndim
        ) + np.expand_dims(mean_used, axis=mean_used.ndim)
        out = np.transpose(out, axes=(out.ndim - 1, *range(out.ndim - 1)))
        return out.squeeze().astype(F32)


# --------------------------------------------------------------------------------------------
# L3 driver            decomposition.py:643-909
# --------------------------------------------------------------------------------------------
@dataclass
class OracleResult:
    u: sp.csr_matrix
    r: np.ndarray
    s: np.ndarray
    vt: np.ndarray
    mean_img: np.ndarray
    std_img: np.ndarray
    shape: Tuple[int, int, int]
    order: str
    ranks: np.ndarray  # per block, (dim1 outer, dim2 inner)
    block_starts: List[Tuple[int, int]]
    thresholds: Tuple[float, float]
    spatial_basis: np.ndarray
    mixing: np.ndarray  # P
    v_init: np.ndarray  # v_cropped (R, t)
    v_full: np.ndarray  # P^T U^T Yc  (k, T)
    block_diags: list = field(default_factory=list)

    def to_pmdarray(self) -> PMDArrayOracle:
        return PMDArrayOracle(self.u, self.r, self.s, self.vt, self.shape, self.order, self.mean_img, self.std_img)


def localmd_decomposition_oracle(
    movie,
    block_sizes,
    frame_range: int,
    draws: Draws,
    max_components: int = 50,
    background_rank: int = 15,
    sim_conf: int = 5,
    frame_batch_size: int = 10000,
    max_consecutive_failures: int = 1,
    rank_prune: bool = False,
    rank_prune_factor: float = 0.33,
    temporal_avg_factor: int = 10,
    spatial_avg_factor: int = 2,
    order: str = "F",
    window_chunks: Optional[int] = None,
    compute_normalizer: bool = True,
    pixel_weighting: Optional[np.ndarray] = None,
    spatial_denoiser: Optional[Callable] = None,
    temporal_denoiser: Optional[Callable] = None,
    timings: Optional[dict] = None,
) -> OracleResult:
    """decomposition.py:643-909 with every random draw taken from `draws`."""
    import time

    def tick(name, t0):
        if timings is not None:
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0

    T, d1, d2 = movie.shape
    check_fov_size((d1, d2))
    t0 = time.perf_counter()
    mean_img, std_img = mean_and_noise(movie, compute_normalizer)
    tick("stats", t0)
    t0 = time.perf_counter()
    spatial_basis = background_basis(movie, mean_img, std_img, draws.bg_frames, draws.bg_sketch, background_rank, order)
    tick("background", t0)

    if window_chunks is None:
        window_chunks = frame_range
    if T < frame_range:
        frame_range = T
        frames = list(range(T))
        if frame_range <= window_chunks:
            window_chunks = frame_range
    else:
        if frame_range <= window_chunks:
            window_chunks = frame_range
        window_chunk_candidates(frame_range, T, window_chunks)  # raises like the reference
        frames = list(draws.init_frames)

    block_sizes = update_block_sizes(block_sizes, (d1, d2))
    bh, bw = block_sizes

    t0 = time.perf_counter()
    if draws.thresholds is not None:
        thr_s, thr_t = draws.thresholds
    else:
        thr_s, thr_t = threshold_heuristic(draws.sim_noise, draws.sim_sketch, sim_conf)
    tick("thresholds", t0)

    t0 = time.perf_counter()
    data, temporal_basis_crop = temporal_crop_with_filter(
        movie, frames, mean_img, std_img, spatial_basis, order, frame_batch_size
    )
    if pixel_weighting is not None:
        data *= pixel_weighting[:, :, None]
    tick("init_filter", t0)

    dim_1_iters = tile_starts(d1, bh)
    dim_2_iters = tile_starts(d2, bw)
    block_weights = pyramid_weights(bh, bw)
    sparse_indices = np.arange(d1 * d2).reshape((d1, d2), order=order)

    if temporal_avg_factor >= data.shape[2]:
        raise ValueError("Need at least {} frames".format(temporal_avg_factor))
    if data.shape[2] // temporal_avg_factor <= max_components:
        max_components = int(data.shape[2] // temporal_avg_factor)
    crop = (data.shape[2] // temporal_avg_factor) * temporal_avg_factor
    temporal_basis_crop = temporal_basis_crop[:, :crop]

    t0 = time.perf_counter()
    rows, cols, vals = [], [], []
    cumulative = np.zeros((d1, d2))
    total_temporal = []
    ranks, starts, bdiags = [], [], []
    col = 0
    bi = 0
    for k in dim_1_iters:
        for j in dim_2_iters:
            subset = data[k : k + bh, j : j + bw, :].astype(F32)[:, :, :crop]
            sc, tc, dg = windowed_pmd(
                window_chunks,
                subset,
                max_components,
                thr_s,
                thr_t,
                max_consecutive_failures,
                temporal_avg_factor,
                spatial_avg_factor,
                draws.block_sketches[bi],
                spatial_denoiser,
                temporal_denoiser,
            )
            bi += 1
            total_temporal.append(tc)
            sc = sc * block_weights[:, :, None]
            cumulative[k : k + bh, j : j + bw] += block_weights
            r_here = sc.shape[2]
            ridx = np.broadcast_to(sparse_indices[k : k + bh, j : j + bw][:, :, None], sc.shape)
            cidx = np.broadcast_to(np.arange(col, col + r_here)[None, None, :], sc.shape)
            rows.append(ridx.ravel())
            cols.append(cidx.ravel())
            vals.append(sc.ravel())
            col += r_here
            ranks.append(r_here)
            starts.append((k, j))
            bdiags.append(dg)
    u_r = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(d1 * d2, col))
    v_cropped = np.concatenate(total_temporal, axis=0)
    wdiag = np.zeros((d1 * d2,))
    wdiag[sparse_indices.flatten(order=order)] = cumulative.flatten(order=order)
    u_r = sp.diags([(1 / wdiag).ravel()], [0]).dot(u_r)
    u_r = sp.hstack([u_r, sp.coo_matrix(spatial_basis)])
    v_cropped = np.concatenate([v_cropped, temporal_basis_crop], axis=0)
    tick("blocks", t0)

    t0 = time.perf_counter()
    if rank_prune:
        if rank_prune_factor <= 0 or rank_prune_factor > 1:
            raise ValueError("Rank prune factor should be a value in the interval (0, 1]")
        min_dim = min(u_r.shape[1], v_cropped.shape[1])
        shape = (v_cropped.shape[1], int(min_dim * rank_prune_factor))
        omega = draws.prune_sketch(shape) if callable(draws.prune_sketch) else draws.prune_sketch
        assert tuple(omega.shape) == shape, (omega.shape, shape)
        reform = np.asarray(v_cropped, dtype=F32) @ np.asarray(omega, dtype=F32)
        p = compute_lowrank_factorized_svd(u_r, reform, only_left=True)
    else:
        p = compute_lowrank_factorized_svd(u_r, v_cropped, only_left=True)
    tick("whiten", t0)

    t0 = time.perf_counter()
    v_full = v_projection(movie, u_r, p, mean_img, std_img, order, frame_batch_size)
    tick("projection", t0)
    t0 = time.perf_counter()
    r, s, vt = projected_svd(p, v_full)
    good = s != 0
    r, s, vt = r[:, good], s[good], vt[good, :]
    tick("final_svd", t0)
    return OracleResult(
        u=sp.csr_matrix(u_r),
        r=r,
        s=s,
        vt=vt,
        mean_img=mean_img,
        std_img=std_img,
        shape=(T, d1, d2),
        order=order,
        ranks=np.array(ranks, dtype=np.int32),
        block_starts=starts,
        thresholds=(float(thr_s), float(thr_t)),
        spatial_basis=spatial_basis,
        mixing=p,
        v_init=np.asarray(v_cropped),
        v_full=v_full,
        block_diags=bdiags,
    )
