"""localmd_decomposition on one B200 (or one frame shard per GPU): the host-side mirror of the
reference driver (decomposition.py:643-909) over the sm_100a kernels in csrc/.

Same signature, defaults, guards and result object as the reference.  Extra keyword-only arguments:
  draws      object with the reference's random quantities (same field names as the test oracle's
             `Draws`): bg_frames, bg_sketch, init_frames, sim_noise, sim_sketch, thresholds,
             block_sketches, prune_sketch.  Any field left None is generated on the device.
  seed       seed for everything generated here.
  device     CUDA device (default: current).
  timings    dict filled with per-stage GPU milliseconds (CUDA events).
  details    dict filled with intermediates (ranks, thresholds, background basis, ...) for tests.
README aliases block_height/block_width/frames_to_init are accepted.
"""
import concurrent.futures
import contextlib
import datetime
import math
import os
import time
import sys
from typing import Callable, Optional

import numpy as np
import scipy.sparse
import torch

from . import ops, sharding
from .dataset import DeviceMovie
from .pmdarray import PMDArray


def display(msg):
    tag = "[" + datetime.datetime.today().strftime("%y-%m-%d %H:%M:%S") + "]: "
    sys.stdout.write(tag + msg + "\n")
    sys.stdout.flush()


# ---------------------------------------------------------------------------------------------
# host logic shared with the reference (guards, tiling, weights, window selection)
# ---------------------------------------------------------------------------------------------
def check_fov_size(fov_dims, min_allowed_value=10):
    """decomposition.py:616-635."""
    for k in fov_dims:
        if k < min_allowed_value:
            raise ValueError(
                "At least one FOV dimension is lower than {}, too small to process".format(min_allowed_value)
            )


def update_block_sizes(blocks, fov_shape, min_block_value=10):
    """decomposition.py:572-613."""
    if blocks[0] < min_block_value or blocks[1] < min_block_value:
        raise ValueError(
            "One of the block dimensions was less than min allowed value of {}, "
            "set to a larger value".format(min_block_value)
        )
    out = []
    for b, n in zip(blocks, fov_shape):
        if b > n:
            display("Blocksize was set to {} but corresponding dimension has size {}. Truncating to {}".format(b, n, n))
        out.append(min(int(b), int(n)))
    return out


def tile_starts(n, b):
    """decomposition.py:698, 723-739."""
    overlap = math.ceil(b / 2)
    it = list(range(0, n - b + 1, b - overlap))
    if it[-1] != n - b and n - b != 0:
        it.append(n - b)
    return it


def pyramid_weights(bh, bw):
    """decomposition.py:742-750 (even block sizes only, as in the reference)."""
    if bh % 2 or bw % 2:
        raise ValueError("block sizes must be even (the reference's weighting, decomposition.py:749, fails otherwise)")
    w = np.ones((bh, bw), dtype=np.float32)
    hbh, hbw = bh // 2, bw // 2
    w[:hbh, :hbw] += np.minimum(np.tile(np.arange(0, hbw), (hbh, 1)), np.tile(np.arange(0, hbh), (hbw, 1)).T)
    w[:hbh, hbw:] = np.fliplr(w[:hbh, :hbw])
    w[hbh:, :] = np.flipud(w[:hbh, :])
    return w


def identify_window_chunks(frame_range, total_frames, window_chunks, rng, starting_points=None):
    """decomposition.py:528-569."""
    if frame_range > total_frames:
        raise ValueError("Requested more frames than available")
    if window_chunks > frame_range:
        raise ValueError("The size of each temporal chunk is bigger than frame range")
    num_intervals = math.ceil(frame_range / window_chunks)
    available = np.arange(0, total_frames, window_chunks)
    if available[-1] > total_frames - window_chunks:
        available[-1] = total_frames - window_chunks
    if starting_points is None:
        starting_points = rng.choice(available, size=num_intervals, replace=False)
    starting_points = np.sort(np.asarray(starting_points))
    frames = []
    for k in starting_points:
        frames.extend(range(int(k), int(min(k + window_chunks, total_frames))))
    return frames


_ACTIVE_TIMER = None
_SIDE_STREAMS = {}
_HOST_POOL = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="pmd-host-tables")


@contextlib.contextmanager
def _short_switch_interval(seconds=1e-4):
    """The launching thread and the host-table worker thread hand the interpreter lock back and forth (every ctypes call
    into the library releases it).  With CPython's default 5 ms switch interval the thread that wants the lock back can
    wait that long while the other runs Python code: measured as occasional 5-7 ms holes in the device queue where the
    driver waits for the worker's tables (whitening, projection).  The interval is restored on exit."""
    old = sys.getswitchinterval()
    sys.setswitchinterval(seconds)
    try:
        yield
    finally:
        sys.setswitchinterval(old)


def _side_stream(dev, priority=0):
    """One cached side stream per (device, priority): the caching allocator keeps a pool per stream, so a fresh stream per
    call would cudaMalloc (and synchronise) every time."""
    key = (str(dev), priority)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=priority)
    return _SIDE_STREAMS[key]


_HOST_TRACE = [] if os.environ.get("PMD_HOST_TRACE") else None   # development: host wall-clock marks of one job


def _ht(label):
    if _HOST_TRACE is not None:
        _HOST_TRACE.append((label, time.perf_counter()))


def _submark(name):
    """Sub-stage CUDA-event mark (only when the caller asked for timings with detail)."""
    if _ACTIVE_TIMER is not None and _ACTIVE_TIMER.detail:
        _ACTIVE_TIMER.mark(name)


class _Timer:
    def __init__(self, timings, device):
        self.timings, self.device, self.events = timings, device, []
        self.detail = bool(timings is not None and timings.get("__detail__", False))

    def mark(self, name):
        if self.timings is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        self.events.append((name, ev))

    def finish(self):
        if self.timings is None or len(self.events) < 2:
            return
        torch.cuda.synchronize(self.device)
        for (_, a), (name, b) in zip(self.events[:-1], self.events[1:]):
            dt = a.elapsed_time(b)
            self.timings[name] = self.timings.get(name, 0.0) + dt
            if self.detail and "." in name:  # sub-stage marks also count towards their stage
                top = name.split(".")[0]
                self.timings[top] = self.timings.get(top, 0.0) + dt


def _as_dev(x, device, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x))).to(device=device, dtype=dtype).contiguous()


def sym_eigh_desc_abs(g):
    """Eigen-decomposition of a symmetric matrix ordered by |lambda| descending -- what
    jnp.linalg.svd(hermitian=True) returns (decomposition.py:984, 1090, 1129) -- in float64.
    Small problems use the Jacobi kernel; larger ones currently go through torch.linalg.eigh (cuSOLVER).
    Returns (|lambda| float64 (n,), vectors float64 (n,n))."""
    n = g.shape[0]
    g = g.to(torch.float64)
    if n <= 112:
        w, vecs = ops.jacobi_eigh(g.contiguous().clone()[None], mode=0)
        w, vecs = w[0], vecs[0].to(torch.float64)
    else:
        w, vecs = torch.linalg.eigh(g)
    order = torch.argsort(w.abs(), descending=True, stable=True)
    return w.abs()[order], vecs[:, order].contiguous()


# ---------------------------------------------------------------------------------------------
# stage functions
# ---------------------------------------------------------------------------------------------
def compute_mean_and_noise(movie: DeviceMovie, compute_normalizer=True, group=None):
    """K1 over this rank's frame shard (+ all-reduce when `group` is a process group).
    pmd_loader.py:203-291."""
    dev = movie.device
    d = movie.d
    mean = torch.zeros(d, dtype=torch.float32, device=dev)
    noise = torch.zeros(d, dtype=torch.float32, device=dev)
    n_var = 0
    for _, chunk in movie.batches():
        mp, npart, nv = ops.stats_pass(chunk, movie.T_total)
        mean += mp.sum(dim=0)
        noise += npart.sum(dim=0)
        n_var += nv
    if group is not None:
        import torch.distributed as dist

        cnt = torch.tensor([n_var], dtype=torch.int64, device=dev)
        dist.all_reduce(mean, group=group)
        dist.all_reduce(noise, group=group)
        dist.all_reduce(cnt, group=group)
        n_var = int(cnt.item())
    flag = bool(compute_normalizer) and movie.T_total >= 256
    if flag and n_var > 0:
        std = noise / float(n_var)
        std = torch.where(std == 0, torch.ones_like(std), std)
    else:
        std = torch.ones(d, dtype=torch.float32, device=dev)
    return mean, std


def background_basis(movie: DeviceMovie, mean, std, bg_frames, bg_sketch, background_rank, group=None, bounds=None):
    """pmd_loader.py:300-314 + 46-68: rSVD of <= 1000 standardised frames -> (K, d) orthonormal rows.
    With a process group the sampled frames are fetched from their owner ranks; every rank computes the same basis."""
    dev = movie.device
    l = bg_sketch.shape[1]
    if l > 32:   # wider sketches than the streaming kernels take (background_rank > 22): library contractions
        raw = sharding.gather_frames(movie, bg_frames, group, bounds)
        a_t = ops.standardize_frames(raw, torch.arange(raw.shape[0], device=dev), mean, std)  # (n, d) = A^T
        with ops.fp32_matmul():
            y = torch.matmul(a_t.t(), bg_sketch).contiguous()  # (d, l)
            q = ops.orthonormalize_cols(y[None])[0]  # (d, l)
            bmat = torch.matmul(q.t(), a_t.t()).contiguous()  # (l, n)
            _, e = ops.jacobi_eigh(ops.gram_rows(bmat[None]), mode=0)
            u = torch.matmul(q, e[0][:, :background_rank])  # (d, K)
        return u.t().contiguous()
    # the sampled frames, standardised, in the pixel-major layout (d, ld): both contractions of the rSVD stream it once
    n = len(bg_frames)
    if group is None:
        movie2d, idx = movie.frame_source(bg_frames)   # the resident movie itself + row indices: no gathered copy
    else:
        movie2d, idx = sharding.gather_frames(movie, bg_frames, group, bounds), torch.arange(n, device=dev)
    yt = ops.standardize_frames_t(movie2d, idx, mean, std, ld=(n + 31) // 32 * 32)   # (d, ld), 128-byte aligned rows
    y = ops.rows_sketch(yt, n, bg_sketch.contiguous())                # (d, l) = A Omega               (pmd_loader.py:58)
    for _ in range(2):                                                # orthonormal basis of its range (pmd_loader.py:59)
        y = ops.rows_times_small(y[None], ops.chol_whiten(ops.gram_cols(y[None], l)))[0]   # CholQR2, float64 Gram
    qt = ops.rows_times_small(y[None], torch.eye(l, dtype=torch.float32, device=dev)[None], transposed=True)[0]   # (l, d) = Q^T
    bmat = ops.bg_project_t(yt, qt, n_ranges=296)[:, :n].contiguous()  # (l, n) = Q^T A                (pmd_loader.py:60)
    _, e = ops.jacobi_eigh(ops.gram_rows(bmat[None]), mode=0)
    # left singular vectors of A restricted to the leading `background_rank` (pmd_loader.py:61-68), as rows (K, d)
    return ops.rows_times_small(y[None], e[:, :, :background_rank].contiguous(), transposed=True)[0]


def simulate_thresholds(bh, bw, t_win, sim_conf, draws, gen, device, iters=250, chunk=None, defer=False):
    """decomposition.py:147-189: roughness statistics of the rank-1 rSVD of pure-noise blocks.
    Every simulated block is its own tiny pixel-major movie (b, ld) for the block kernels.
    defer=True: the device work is only enqueued; the returned callable fetches the statistics and takes the percentiles."""
    b = bh * bw
    ld = (t_win + 3) // 4 * 4
    if chunk is None:   # as many simulated blocks per batch as ~3 GB of noise allow (all 250 at C2: fewer, larger launches)
        chunk = int(max(10, min(iters, (3 << 30) // (4 * b * ld))))
    starts = torch.zeros((chunk, 2), dtype=torch.int32, device=device)
    sp, tp = [], []
    have_noise = getattr(draws, "sim_noise", None) is not None
    have_sk = getattr(draws, "sim_sketch", None) is not None
    n_iters = len(draws.sim_noise) if have_noise else iters
    for s0 in range(0, n_iters, chunk):
        m = min(chunk, n_iters - s0)
        if not have_noise and ld == t_win:
            # drawn in place: zero fill + a temporary + a strided copy would move 8 GB instead of 2 GB at C2
            noise = torch.empty((m, b, ld), dtype=torch.float32, device=device).normal_(generator=gen)
        else:
            noise = torch.zeros((m, b, ld), dtype=torch.float32, device=device)
            if have_noise:
                for i in range(m):
                    noise[i, :, :t_win] = _as_dev(np.asarray(draws.sim_noise[s0 + i]).reshape(b, t_win), device)
            else:
                noise[:, :, :t_win] = torch.randn((m, b, t_win), generator=gen, device=device, dtype=torch.float32)
        if have_sk:
            sk = torch.stack([_as_dev(draws.sim_sketch[s0 + i], device) for i in range(m)])
        else:
            sk = torch.randn((m, t_win, 11), generator=gen, device=device, dtype=torch.float32)
        l = sk.shape[2]
        y = torch.bmm(noise[:, :, :t_win], sk)  # (m, b, l)
        q = ops.block_orth(y) if ops.block_orth_fits(b, l) else ops.orthonormalize_cols(y)  # (m, b, l)
        rp = (l + 3) // 4 * 4
        qp = torch.zeros((m, b, rp), dtype=torch.float32, device=device)
        qp[:, :, :l] = q
        bq = ops.block_project_tc(noise, b * ld, ld, bw, starts[:m], bh, bw, qp, l)  # (m, l, ld)
        _, e = ops.jacobi_eigh(ops.gram_rows(bq), mode=0)
        e1 = e[:, :, :1].contiguous()  # (m, l, 1)
        u1 = torch.zeros((m, b, 4), dtype=torch.float32, device=device)
        u1[:, :, :1] = torch.bmm(q, e1)
        v1 = torch.bmm(e1.transpose(1, 2), bq).contiguous()  # (m, 1, ld) == s * v
        ss, ts, _ = ops.block_stats_rank(u1, v1, bh, bw, 1, float("inf"), float("inf"), 1, t=t_win)
        sp.append(ss.reshape(-1))
        tp.append(ts.reshape(-1))
    sp, tp = torch.cat(sp), torch.cat(tp)

    if not defer:
        return np.percentile(sp.cpu().numpy().flatten(), sim_conf), np.percentile(tp.cpu().numpy().flatten(), sim_conf)
    # deferred: the statistics travel to page-locked host memory on the CURRENT (side) stream; the percentiles are taken
    # when the thresholds are first needed -- at the rank decision at the end of the block stage, by which time the host has
    # long run ahead of the device, so neither the copy nor the host-side percentile stalls the main stream
    host = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in (sp, tp)]
    for h, x in zip(host, (sp, tp)):
        h.copy_(x, non_blocking=True)
    done = torch.cuda.Event()
    done.record(torch.cuda.current_stream(device))
    memo = []

    def finish():
        if not memo:
            done.synchronize()
            memo.append((float(np.percentile(host[0].numpy().flatten(), sim_conf)), float(np.percentile(host[1].numpy().flatten(), sim_conf))))
        return memo[0]

    return finish


def block_decompositions(yt, t, d2, starts_dev, bh, bw, r, taf, saf, thr_s, thr_t, mcf, sketches, spatial_denoiser=None,
                         temporal_denoiser=None):
    """single_block_md (decomposition.py:235-330) for all blocks at once.
    temporal_denoiser / spatial_denoiser: the reference's hooks (decomposition.py:300, 310) as torch callables on CUDA
    tensors, (n_traces, t) -> (n_traces, t) and (n_images, bh, bw) -> (n_images, bh, bw), applied to all blocks at once.
    yt (d, ld) float32: pixel-major standardised init movie, t frames (a multiple of taf), zero padded
    to ld.  sketches (nb, t//taf, r+10).
    Returns U (nb, b, rp), V (nb, r, ld), ranks (nb,), sstat, tstat."""
    dev = yt.device
    d, ld = yt.shape
    nb = starts_dev.shape[0]
    rp = (r + 3) // 4 * 4
    ph, pw = -(-bh // saf), -(-bw // saf)
    pooled = None
    d1 = d // d2
    fov_pool = None
    if bh % saf == 0 and bw % saf == 0 and d1 % saf == 0 and d2 % saf == 0 and d1 // saf <= 65535 and bool(
            (starts_dev % saf == 0).all()):
        # Every block starts on the pooling grid: its pooled block is a window of the POOLED FIELD OF VIEW, so the movie
        # is pooled once (the pooling entry point applied to non-overlapping saf x d2 row bands: same arithmetic, bit
        # identical cells) instead of once per overlapping block (4x the reads, 4x the writes at stride b/2)
        d1p, d2p = d1 // saf, d2 // saf
        bands = torch.stack([torch.arange(0, d1, saf, dtype=torch.int32, device=dev),
                             torch.zeros(d1p, dtype=torch.int32, device=dev)], dim=1).contiguous()
        yp, ypa = ops.block_pool_full(yt, t, d2, bands, saf, d2, saf, taf)      # (d1p, d2p, ld), (d1p, d2p, t')
        sp = (starts_dev // saf).to(torch.int32).contiguous()
        qp = torch.arange(ph * pw, device=dev)
        pixp = (sp[:, 0:1].to(torch.int64) + (qp // pw)[None, :]) * d2p + sp[:, 1:2].to(torch.int64) + (qp % pw)[None, :]
        bta = ypa.reshape(d1p * d2p, -1)[pixp].contiguous()                      # (nb, P, t')
        fov_pool = (yp.reshape(d1p * d2p, ld), d2p, sp)
        del ypa
    elif nb * ph * pw * ld * 4 <= 24 << 30:
        # B_ds at full time resolution is kept (5.2 GB at C2) so that U_ds^T B_ds contracts over the pooled pixels
        pooled, bta = ops.block_pool_full(yt, t, d2, starts_dev, bh, bw, saf, taf)  # (nb, P, ld), (nb, P, t')
    else:
        bta = ops.block_pool_tavg(yt, t, d2, starts_dev, bh, bw, saf, taf)  # (nb, P, t') = B_ta
    _submark("blocks.pool")
    P = bta.shape[1]
    l = sketches.shape[2]
    if P > l:
        tq = bta.shape[2]
        on_tc = l <= 64 and l % 4 == 0 and tq % 4 == 0 and pw % 2 == 0
        if on_tc:
            # both sketch products on the tensor cores through the block kernels: B_ta[b] (P pooled pixels x t' frames) is a
            # ph x pw "block" of a pixel-major movie with batch stride P t' (0.15 + 0.1 ms against 0.4 + 0.4 ms for the
            # library's SIMT batched GEMMs at C2)
            zero_starts = torch.zeros((nb, 2), dtype=torch.int32, device=dev)
            y = ops.block_spatial_tc(bta, P * tq, tq, pw, zero_starts, ph, pw, sketches.transpose(1, 2).contiguous(), l)  # (nb, P, l) = B_ta Omega
        else:
            y = torch.bmm(bta, sketches)  # (nb, P, l)
        q = ops.block_orth(y) if ops.block_orth_fits(P, l) else ops.orthonormalize_cols(y)
        if on_tc:
            bq = ops.block_project_tc(bta, P * tq, tq, pw, zero_starts, ph, pw, q, l)  # (nb, l, t') = Q^T B_ta
        else:
            bq = torch.bmm(q.transpose(1, 2), bta).contiguous()  # (nb, l, t')
        # sketch-stage SVD (decomposition.py:66): float32 rotations on the float64 Gram, the accuracy of the reference's
        # own float32 SVD; its output only seeds the temporal basis that the full-resolution steps below refine
        _, e = ops.jacobi_eigh(ops.gram_rows(bq), mode=0, sweeps_f32=True)
        uds = torch.bmm(q, e[:, :, :r]).contiguous()  # (nb, P, r)
        del y, q, bq, e
    else:
        # reduced QR of a P x l sketch with P <= l spans R^P: the rSVD is the exact SVD of B_ta
        if r > P:
            raise TypeError("max_components larger than the pooled block (jax.lax.dynamic_slice would fail)")
        _, e = ops.jacobi_eigh(ops.gram_rows(bta), mode=0)
        uds = e[:, :, :r].contiguous()
    del bta
    _submark("blocks.rsvd")
    if fov_pool is not None:
        udp = torch.zeros((nb, ph * pw, rp), dtype=torch.float32, device=dev)
        udp[:, :, :r] = uds
        ypm, d2p, sp = fov_pool
        vds = ops.block_project_tc(ypm, 0, ld, d2p, sp, ph, pw, udp, r)  # (nb, r, ld) = U_ds^T B_ds
        del fov_pool, ypm, udp
    elif pooled is not None:
        udp = torch.zeros((nb, ph * pw, rp), dtype=torch.float32, device=dev)
        udp[:, :, :r] = uds
        zero_starts = torch.zeros((nb, 2), dtype=torch.int32, device=dev)
        vds = ops.block_project_tc(pooled, ph * pw * ld, ld, pw, zero_starts, ph, pw, udp, r)  # (nb, r, ld) = U_ds^T B_ds
        del pooled, udp
    else:
        w4 = ops.block_unpool(uds, bh, bw, saf, rp)  # (nb, b, rp): w4^T B == U_ds^T B_ds
        vds = ops.block_project_tc(yt, 0, ld, d2, starts_dev, bh, bw, w4, r)  # (nb, r, ld)
        del w4
    _submark("blocks.project1")
    if temporal_denoiser is not None:
        # decomposition.py:300: the sketch traces of every block are denoised before they define the temporal basis
        tr = temporal_denoiser(vds[:, :, :t].reshape(nb * r, t).contiguous())
        if tuple(tr.shape) != (nb * r, t):
            raise ValueError("temporal_denoiser must return a (n_traces, n_frames) tensor of the shape it was given")
        vds = torch.zeros((nb, r, ld), dtype=torch.float32, device=dev)
        vds[:, :, :t] = tr.to(torch.float32).reshape(nb, r, t)
    if spatial_denoiser is not None:
        # decomposition.py:301-313: the reference denoises the images S = B Vb^T with Vb = the right singular vectors
        # of the sketch traces; a nonlinear denoiser depends on those specific images (not only on their span), so Vb
        # is formed explicitly: V_ds V_ds^T = E diag(w) E^T, Vb = diag(w)^-1/2 E^T V_ds (rows by descending w)
        w, e = ops.jacobi_eigh(ops.gram_rows(vds), mode=0)
        scale = torch.where(w > 0, w.clamp_min(1e-300).rsqrt(), torch.zeros_like(w)).to(torch.float32)
        vds = torch.bmm((e * scale[:, None, :]).transpose(1, 2).contiguous(), vds)
    # S' = B V_ds^T has the same column space as the reference's S = B Vb^T (Vb = orthonormalised rows of V_ds,
    # decomposition.py:301-306), and only that space is used downstream (Uf = orth(S), 315); the rows of V_ds are the
    # nearly orthogonal sketch directions, so S' only needs the column scaling that CholQR's relative pivots apply anyway.
    s_raw = ops.block_spatial_tc(yt, 0, ld, d2, starts_dev, bh, bw, vds, rp)  # (nb, b, rp): tcgen05, 3xTF32
    del vds
    if spatial_denoiser is not None:
        # decomposition.py:307-313: every (bh, bw) image of the spatial projection is denoised
        imgs = s_raw[:, :, :r].reshape(nb, bh, bw, r).permute(0, 3, 1, 2).reshape(nb * r, bh, bw).contiguous()
        den = spatial_denoiser(imgs)
        if tuple(den.shape) != (nb * r, bh, bw):
            raise ValueError("spatial_denoiser must return a (n_images, block height, block width) tensor of the shape it was given")
        s_raw = torch.zeros((nb, bh * bw, rp), dtype=torch.float32, device=dev)
        s_raw[:, :, :r] = den.to(torch.float32).reshape(nb, r, bh * bw).transpose(1, 2)
    _submark("blocks.spatial")
    if ops.block_orth_fits(bh * bw, r):
        uf = ops.block_orth(s_raw, r)  # (nb, b, rp)
    else:  # large blocks: Gram/Jacobi orthonormalisation in global memory
        uf = ops.orthonormalize_cols(s_raw, r)
    del s_raw
    _submark("blocks.orth_s")
    vn = ops.block_project_tc(yt, 0, ld, d2, starts_dev, bh, bw, uf, r)  # (nb, r, ld): tcgen05, 3xTF32
    _submark("blocks.project2")
    g6 = ops.gram_rows(vn)
    _submark("blocks.gram2")
    _, lmat = ops.jacobi_eigh(g6, mode=0)  # (nb, r, r)
    _submark("blocks.jacobi2")
    lpad = torch.zeros((nb, rp, rp), dtype=torch.float32, device=dev)
    lpad[:, :r, :r] = lmat
    # rotation into the singular vectors (decomposition.py:319-323): V <- L^T V on the tensor cores (_rows_times_coef_t); the
    # small U <- U L stays a library batched GEMM pinned to full float32 (0.27 ms; own kernel measured 0.57 ms)
    with ops.fp32_matmul():
        u = torch.bmm(uf, lpad)  # (nb, b, rp)
    v = _rows_times_coef_t(lpad[:, :r, :], vn, r)  # (nb, r, ld)
    del uf, vn
    _submark("blocks.bmm_uv")
    if callable(thr_s):   # deferred threshold simulation (see simulate_thresholds): resolved here, where it is first needed
        thr_s, thr_t = thr_s()
    sstat, tstat, ranks = ops.block_stats_rank(u, v, bh, bw, r, thr_s, thr_t, mcf, t=t)
    _submark("blocks.stats")
    return u, v, ranks, sstat, tstat


def _rows_times_coef_t(coef, v, r):
    """out[b] = coef[b][:, :r]^T v[b]  for coef (nb, r, rp) (zero padded columns) and v (nb, r, ld): the r x r mixing of every
    block's temporal rows.  Runs on the tensor cores through the block projection kernel -- v[b] is a 1 x r "pixel block"
    of a pixel-major movie whose coefficient images are the columns of coef[b] (1.16 ms against 2.25 ms for the library's
    SIMT batched GEMM at C2, scripts/debug/rot_bench.py); odd r (no even-width TMA box) falls back to the library."""
    nb, _, ld = v.shape
    if r % 2 == 0:
        return ops.block_project_tc(v, r * ld, ld, r, torch.zeros((nb, 2), dtype=torch.int32, device=v.device), 1, r,
                                    coef.contiguous(), r)
    with ops.fp32_matmul():
        return torch.bmm(coef[:, :, :r].transpose(1, 2), v)


def _append_components(final, counter, comps, n_new):
    """final[b, :, counter[b] : counter[b] + n_new[b]] = comps[b, :, : n_new[b]]  (ragged append, all blocks at once)."""
    nb, bpix, rp = final.shape
    j = torch.arange(comps.shape[2], device=final.device)[None, :]
    valid = j < n_new[:, None]
    if not bool(valid.any()):
        return
    blk = torch.arange(nb, device=final.device)[:, None].expand_as(valid)[valid]
    src = j.expand_as(valid)[valid]
    dst = (counter[:, None] + j)[valid]
    ft, ct = final.transpose(1, 2), comps.transpose(1, 2)   # (nb, comps, pixels) views
    ft[blk, dst] = ct[blk, src]


def block_decompositions_windowed(yt, t, d2, starts_dev, bh, bw, r, taf, saf, thr_s, thr_t, mcf, sketches, window,
                                  spatial_denoiser=None, temporal_denoiser=None):
    """windowed_pmd (decomposition.py:410-525) for all blocks at once, window_chunks < frame_range.
    The first window (and any later window of a block that has kept nothing yet) is fitted by single_block_md
    (235-330); later windows fit the residual after projecting out the components kept so far
    (single_residual_block_md, 333-387: time average only, no spatial pooling); a block stops once it holds r
    components.  sketches: list over windows of (nb, window // taf, r + 10).  Same return values as
    block_decompositions, with V = (kept U)^T block over all t frames (get_temporal_projector, 390-407)."""
    if callable(thr_s):
        thr_s, thr_t = thr_s()
    dev = yt.device
    d, ld = yt.shape
    nb = starts_dev.shape[0]
    rp = (r + 3) // 4 * 4
    bpix = bh * bw
    if window % taf:
        raise ValueError("window_chunks must be a multiple of temporal_avg_factor")
    start_points = list(range(0, t, window))
    if start_points and start_points[-1] + window > t:
        start_points[-1] = t - window
    ldw = (window + 3) // 4 * 4
    final = torch.zeros((nb, bpix, rp), dtype=torch.float32, device=dev)
    counter = torch.zeros((nb,), dtype=torch.int64, device=dev)
    sstat = torch.zeros((nb, r), dtype=torch.float32, device=dev)
    tstat = torch.zeros((nb, r), dtype=torch.float32, device=dev)
    for wi, k in enumerate(start_points):
        active = counter < r
        if not bool(active.any()):
            break
        yw = torch.zeros((d, ldw), dtype=torch.float32, device=dev)
        yw[:, :window] = yt[:, k : k + window]
        sk = sketches[wi]
        fresh = active & (counter == 0) if k != 0 else active
        resid = active & ~fresh
        ia = torch.nonzero(fresh).reshape(-1)
        if ia.numel():
            u_a, _, rk_a, ss_a, ts_a = block_decompositions(
                yw, window, d2, starts_dev[ia].contiguous(), bh, bw, r, taf, saf, thr_s, thr_t, mcf, sk[ia].contiguous(),
                spatial_denoiser, temporal_denoiser)
            n_new = torch.minimum(rk_a.to(torch.int64), r - counter[ia])
            fa, ca = final[ia], counter[ia]
            _append_components(fa, ca, u_a, n_new)
            final[ia] = fa
            counter[ia] = ca + n_new
            if wi == 0:
                sstat[ia], tstat[ia] = ss_a, ts_a
            del u_a
        ib = torch.nonzero(resid).reshape(-1)
        if ib.numel():
            # single_residual_block_md (decomposition.py:333-387) for all residual blocks at once, WITHOUT materialising the
            # residual blocks: with E = the components kept so far (orthonormal columns) and W = E^T block,
            #   time average of the residual  = tavg(block) - E tavg(W)                (averaging is linear)
            #   U_b^T residual                = U_b^T block - (U_b^T E) W
            # so the movie is only touched by the block kernels (pooling entry point with a 1 x 1 spatial window, tensor-core
            # block projection); what remains are products of per-block r x r / bpix x r matrices.
            n = int(ib.numel())
            st_b = starts_dev[ib].contiguous()
            e = final[ib].contiguous()                                               # (n, bpix, rp), columns >= counter are zero
            w_e = ops.block_project_tc(yw, 0, ldw, d2, st_b, bh, bw, e, r)            # (n, r, ldw) = E^T block
            avg = ops.block_pool_tavg(yw, window, d2, st_b, bh, bw, 1, taf)           # (n, bpix, t') = tavg(block)
            we_avg = w_e[:, :, :window].reshape(n, r, window // taf, taf).mean(dim=3)
            skc = sk[ib].contiguous()
            l = skc.shape[2]
            with ops.fp32_matmul():
                avg.baddbmm_(e[:, :, :r], we_avg, alpha=-1.0)                         # decomposition.py:362-366
                if bpix > l:
                    y = torch.bmm(avg, skc)
                    q = ops.block_orth(y) if ops.block_orth_fits(bpix, l) else ops.orthonormalize_cols(y)
                    bq = torch.bmm(q.transpose(1, 2), avg).contiguous()
                    _, ev = ops.jacobi_eigh(ops.gram_rows(bq), mode=0)
                    u_b = torch.bmm(q, ev[:, :, :r])                                  # (n, bpix, r)
                    del y, q, bq
                else:
                    if r > bpix:
                        raise TypeError("max_components larger than the block (jax.lax.dynamic_slice would fail)")
                    _, ev = ops.jacobi_eigh(ops.gram_rows(avg.contiguous()), mode=0)
                    u_b = ev[:, :, :r]
                u_pad = torch.zeros((n, bpix, rp), dtype=torch.float32, device=dev)
                u_pad[:, :, :r] = u_b
                coef = torch.bmm(e[:, :, :r].transpose(1, 2), u_pad)                  # (n, r, rp) = E^T U_b
            v_b = ops.block_project_tc(yw, 0, ldw, d2, st_b, bh, bw, u_pad, r)        # (n, r, ldw) = U_b^T block
            v_b -= _rows_times_coef_t(coef, w_e, r)                                   # ... - (U_b^T E) W
            _, _, rk_b = ops.block_stats_rank(u_pad, v_b, bh, bw, r, thr_s, thr_t, mcf, t=window)
            n_new = torch.minimum(rk_b.to(torch.int64), r - counter[ib])
            fb, cb = final[ib], counter[ib]
            _append_components(fb, cb, u_pad, n_new)
            final[ib] = fb
            counter[ib] = cb + n_new
            del avg, u_b, v_b, u_pad, w_e, e
        del yw
    v = ops.block_project_tc(yt, 0, ld, d2, starts_dev, bh, bw, final, r)   # (nb, r, ld)
    return final, v, counter.to(torch.int32), sstat, tstat


class SparseU:
    """Device form of the sparse spatial matrix: block-component values + dense background rows."""

    def __init__(self, starts, starts_dev, bh, bw, d1, d2, ranks_host, ranks_dev, uvals64, uvals32, bg):
        self.starts, self.starts_dev = starts, starts_dev
        self.bh, self.bw, self.d1, self.d2 = bh, bw, d1, d2
        self.ranks_host, self.ranks_dev = ranks_host, ranks_dev
        self.col0_host = np.concatenate([[0], np.cumsum(ranks_host)[:-1]]).astype(np.int64)
        self.col0_dev = ops.h2d(self.col0_host, ranks_dev.device)
        self.uvals64, self.uvals32 = uvals64, uvals32
        self.bg = bg  # (K, d) float32
        self.n_local = int(ranks_host.sum())
        self.n_cols = self.n_local + bg.shape[0]
        self._tasks = None
        self._csr = None
        self.supertiles = None
        self.strips = None
        st64 = np.asarray(starts, dtype=np.int64).reshape(-1, 2)
        rows = np.unique(st64[:, 0]).tolist() if len(ranks_host) else []
        cols = np.unique(st64[:, 1]).tolist() if len(ranks_host) else []
        regular = len(ranks_host) > 0 and len(rows) * len(cols) == len(ranks_host) and np.array_equal(
            st64, np.stack(np.meshgrid(rows, cols, indexing="ij"), axis=-1).reshape(-1, 2))
        self.strips_tc = None
        self.strips_ts = None
        self._regular = (rows, cols) if regular else None
        self._tc_host = None
        self._ts_host = None
        self._ts_future = None
        self._utu_host = None
        self._utu_tiles = None
        self._tables_started = False
        which = os.environ.get("PMD_K7", "ts")   # development switch between the generations of the projection kernel
        self._want_ts = bool(regular and which == "ts")
        if self._want_ts:
            self._ts_host = True   # placeholder until the first projection call collects the result
        if regular and self._ts_host is None and which != "simt":
            # K7 on the tensor cores (csrc/project_tc.cu): host tables now (native library), device tables and
            # coefficient images at the first projection call
            self._tc_host = ops.make_strips_tc(rows, cols, bh, bw, d1, d2, ranks_host, self.col0_host, bg.shape[0])
        if regular and self._tc_host is None and self._ts_host is None:
            self._build_simt_strips()
        if self.strips is None and self._tc_host is None and self._ts_host is None:
            self._build_supertiles()

    def start_host_tables(self):
        """Host-side tables that depend on the kept ranks (idempotent): the block-pair bookkeeping of U^T U (whitening) inline,
        the strip tables of K7 (3 ms, native routine of the library, releases the GIL) on the worker thread.  The driver calls
        this AFTER it has enqueued the prune-sketch GEMMs, so that the device has work queued meanwhile."""
        if self._tables_started:
            return
        self._tables_started = True
        if self.n_local > 0:
            # 0.2 ms of (mostly native) work once the block-pair list is memoised: done inline -- on the worker thread the
            # result arrived 4.4 ms after it was requested (the worker has to win the interpreter lock from the launching
            # thread first), and the device idled for 3 ms waiting for it
            self._utu_host = ops.utu_host_tables(self.starts, self.bh, self.bw, self.ranks_host)
        if self._want_ts:
            # K7 with TMA-fed raw tiles and the movie operand in tensor memory (csrc/project_ts.cu): host tables on the worker,
            # device tables and coefficient images at the first projection call
            rows, cols = self._regular
            self._ts_future = _HOST_POOL.submit(ops.make_strips_ts, rows, cols, self.bh, self.bw, self.d1, self.d2,
                                                self.ranks_host.copy(), self.col0_host.copy(), self.bg.shape[0])

    def _build_supertiles(self):
        """Tables of the supertile kernel (irregular block lists, geometries neither strip kernel supports)."""
        rows = sorted(set(int(x) for x in self.starts[:, 0])) if len(self.ranks_host) else []
        cols = sorted(set(int(x) for x in self.starts[:, 1])) if len(self.ranks_host) else []
        if len(self.ranks_host) and self.bh * self.bw <= 512 and len(rows) * len(cols) == len(self.ranks_host):
            st = ops.make_supertiles(rows, cols, self.bh, self.bw, self.ranks_host, self.col0_host)
            if st["max_h"] * st["max_w"] <= 2048:
                self.supertiles = {k: (torch.from_numpy(v).to(self.ranks_dev.device) if isinstance(v, np.ndarray) else v)
                                   for k, v in st.items()}

    def _finish_tc(self):
        """Upload the tensor-core tables and build the coefficient images (once, at the first projection call)."""
        st = self._tc_host
        if st is None:
            return
        self._tc_host = None
        dev = self.ranks_dev.device
        self.strips_tc = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
        self.bimg = ops.pack_strips_tc(self.strips_tc, self.uvals32, self.bg, self.bh * self.bw, self.d2)

    def _finish_ts(self, inv_std):
        """Upload the tables of the TMA / tensor-memory kernel and build its coefficient images with 1 / std folded in
        (once per decomposition, at the first projection call)."""
        if self._ts_future is not None:
            self._ts_host = self._ts_future.result()
            self._ts_future = None
        st = self._ts_host
        if st is not None:
            self._ts_host = None
            dev = self.ranks_dev.device
            self.strips_ts = {k: (ops.h2d(v, dev) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
            self._ts_inv = None
        if self.strips_ts is not None and (getattr(self, "bimg_ts", None) is None or self._ts_inv is not inv_std):
            self.bimg_ts = ops.pack_strips_ts(self.strips_ts, self.uvals32, self.bg, inv_std, self.bh * self.bw, self.d2)
            self._ts_inv = inv_std

    def _build_simt_strips(self):
        """Tables of the SIMT strip-streaming kernel (fallback of the tensor-core path: unaligned movies, tiny FOVs)."""
        rows, cols = self._regular
        st = ops.make_strips(rows, cols, self.bh, self.bw, self.d1, self.d2, self.ranks_host, self.col0_host, self.bg.shape[0])
        if st is not None:
            dev = self.ranks_dev.device
            self.strips = {k: (torch.from_numpy(v).to(dev) if k in ("items", "slot_ptr", "tasks") else v) for k, v in st.items()}
            self.upack = ops.pack_strip_u(st, self.uvals32, self.bg, self.bh * self.bw)

    @property
    def tasks(self):
        if self._tasks is None:
            self._tasks = torch.from_numpy(ops.make_tasks(self.ranks_host)).to(self.ranks_dev.device)
        return self._tasks

    def coo_physical(self):
        """(rows = physical pixel ids, cols, float64 values) of all stored entries, exact zeros dropped
        (scipy's sparse product drops them at decomposition.py:853; coo_matrix(dense) at 929)."""
        dev = self.uvals64.device
        bpix = self.bh * self.bw
        q = torch.arange(bpix, device=dev)
        qi, qj = q // self.bw, q % self.bw
        st = self.starts_dev.to(torch.int64)
        pix_b = (st[:, 0:1] + qi[None]) * self.d2 + st[:, 1:2] + qj[None]  # (nb, bpix)
        blk_of_col = torch.repeat_interleave(torch.arange(len(self.ranks_host), device=dev), self.ranks_dev.to(torch.int64))
        rows = pix_b[blk_of_col].reshape(-1)
        cols = torch.arange(self.n_local, device=dev)[:, None].expand(-1, bpix).reshape(-1)
        vals = self.uvals64.reshape(-1)
        K, d = self.bg.shape
        rows = torch.cat([rows, torch.arange(d, device=dev).repeat(K)])
        cols = torch.cat([cols, (self.n_local + torch.arange(K, device=dev))[:, None].expand(-1, d).reshape(-1)])
        vals = torch.cat([vals, self.bg.to(torch.float64).reshape(-1)])
        keep = vals != 0
        return rows[keep], cols[keep], vals[keep]

    def csr(self, row_ids=None):
        """CSR arrays (indptr int64, indices int32, values float64) with rows relabelled by `row_ids`
        (tensor mapping physical pixel -> row id, a permutation; None = physical order), canonical (sorted).
        The physical-order form is sorted once and cached; a relabelled form is a permutation of its row segments
        (every segment is already sorted by column), built by one scatter instead of a second sort."""
        d = self.d1 * self.d2
        if getattr(self, "_csr64", None) is None:
            rows, cols, vals = self.coo_physical()
            order = torch.argsort(rows * self.n_cols + cols)
            rows, cols, vals = rows[order], cols[order], vals[order]
            counts = torch.bincount(rows, minlength=d)
            indptr = torch.zeros(d + 1, dtype=torch.int64, device=rows.device)
            indptr[1:] = torch.cumsum(counts, 0)
            self._csr64 = (indptr, cols.to(torch.int32), vals, rows, counts)
        indptr, cols32, vals, rows, counts = self._csr64
        if row_ids is None:
            return indptr, cols32, vals
        counts_new = torch.zeros_like(counts)
        counts_new[row_ids] = counts
        indptr_new = torch.zeros(d + 1, dtype=torch.int64, device=rows.device)
        indptr_new[1:] = torch.cumsum(counts_new, 0)
        pos = indptr_new[row_ids[rows]] + (torch.arange(rows.numel(), device=rows.device) - indptr[rows])
        cols_new, vals_new = torch.empty_like(cols32), torch.empty_like(vals)
        cols_new[pos] = cols32
        vals_new[pos] = vals
        return indptr_new, cols_new, vals_new

    def export_csr(self, row_ids):
        """((indptr, indices, values64) with rows relabelled by row_ids, (indptr, indices, values32) over physical rows): on a
        regular block grid written directly by pmd_export_csr (no coordinate list, no sort; the buffers are sliced to the
        true number of entries by finish_export), otherwise through the sorted coordinate form (csr / csr_physical32)."""
        if self._regular is None or len(self.ranks_host) == 0 or self.bg.shape[0] == 0:
            return self.csr(row_ids), self.csr_physical32()
        rows, cols = self._regular
        # origins of the block rows / columns from the device copy of the block list (no upload, no synchronisation)
        rs = self.starts_dev[:: len(cols), 0].contiguous()
        cs = self.starts_dev[: len(cols), 1].contiguous()
        return ops.export_csr(self.uvals64, self.bg, self.d1, self.d2, rs, cs, self.bh, self.bw, self.ranks_dev, self.col0_dev,
                              self.n_local, row_ids)

    @staticmethod
    def finish_export(exported):
        """Slice the worst-case buffers of export_csr to the true entry count (one read-back of indptr[-1])."""
        (ip, ix, v), (ip32, ix32, v32) = exported
        nnz = int(ip32[-1].item())
        if ix.numel() != nnz:
            ix, v, ix32, v32 = ix[:nnz], v[:nnz], ix32[:nnz], v32[:nnz]
        return (ip, ix, v), (ip32, ix32, v32)

    def gram(self):
        """U^T U in float64 as (CSR of the local x local part, C = U^T bg^T (n_cols, K)): the two sparse products
        of decomposition.py:974-981 reduce to applying these to the right factor."""
        if getattr(self, "_gram", None) is None:
            self.start_host_tables()
            dev = self.uvals64.device
            bg64 = self.bg.to(torch.float64).contiguous()
            blk_of_col = torch.repeat_interleave(
                torch.arange(len(self.ranks_host), device=dev, dtype=torch.int32), self.ranks_dev.to(torch.int64),
                output_size=self.n_local).contiguous()
            c = ops.project_cols_f64(bg64, self.d2, self.starts_dev, self.bh, self.bw, blk_of_col, self.col0_dev, self.uvals64,
                                     bg64)  # (n_cols, K)
            if self.n_local > 0:
                host, self._utu_host = self._utu_host, None
                rowptr, cols, vals, self._utu_tiles = ops.utu_local_csr(self.starts, self.starts_dev, self.bh, self.bw, self.ranks_host,
                                                                         self.ranks_dev, self.col0_host, self.col0_dev, self.uvals64,
                                                                         host=host)
                csr = (rowptr, cols, vals)
            else:
                csr = None
            self._gram = (csr, c)
        return self._gram

    def utu_times_f64(self, right64):
        """U^T U right in float64 (right: (n_cols, m))."""
        csr, c = self.gram()
        nl = self.n_local
        right64 = right64.contiguous()
        z = torch.empty_like(right64)
        r_bg = right64[nl:]
        if nl > 0:
            rowptr, cols, vals = csr
            r_loc = right64[:nl].contiguous()
            if getattr(self, "_utu_tiles", None) is not None and len(self.ranks_host) <= 65535:
                # dense block-pair tiles on the FP64 tensor cores: the rows of `right` of a block are read once per pair
                ops.utu_apply_tiles(self._utu_tiles, self.ranks_dev, self.col0_dev, rowptr, vals, r_loc, z[:nl])
            else:
                rows = torch.arange(nl, dtype=torch.int32, device=right64.device)
                z[:nl] = ops.reconstruct_f64(rowptr, cols, vals, r_loc, rows).t()
            z[:nl] += torch.matmul(c[:nl], r_bg)
            z[nl:] = torch.matmul(c[:nl].t(), r_loc)
        else:
            z[nl:] = 0
        z[nl:] += torch.matmul(c[nl:], r_bg)
        return z

    def csr_physical32(self):
        if self._csr is None:
            ip, ix, v = self.csr()
            self._csr = (ip.contiguous(), ix.contiguous(), v.to(torch.float32).contiguous())
        return self._csr

    def apply(self, right):
        """(U right)^T as an (m, d) float32 'movie' (m = right.shape[1])."""
        ip, ix, v = self.csr_physical32()
        d = self.d1 * self.d2
        pix = torch.arange(d, dtype=torch.int32, device=right.device)
        return ops.reconstruct(ip, ix, v, right.contiguous(), pix, None, None)

    def project(self, movie2d, mean, inv_std, z):
        """z[:, :n] (R, ldz) (+)= U^T standardised(movie2d)   (K7a + K7b)."""
        n = movie2d.shape[0]
        self.start_host_tables()
        if self._ts_future is not None:
            self._ts_host = self._ts_future.result()
            self._ts_future = None
            if self._ts_host is None and self._regular is not None:   # geometry the TMA kernel does not take: older kernels
                rows, cols = self._regular
                self._tc_host = ops.make_strips_tc(rows, cols, self.bh, self.bw, self.d1, self.d2, self.ranks_host, self.col0_host,
                                                   self.bg.shape[0])
                if self._tc_host is None:
                    self._build_simt_strips()
                if self.strips is None and self._tc_host is None:
                    self._build_supertiles()
        if (self._ts_host is not None or self.strips_ts is not None) and ops.project_stream_ts_ok(movie2d, self.d2, mean):
            self._finish_ts(inv_std)
            ops.project_stream_ts(movie2d, self.d2, self.strips_ts, self.bimg_ts, mean, z[: self.n_local], z[self.n_local :],
                                  mark=_submark)
            return
        if self._ts_host is not None or self.strips_ts is not None:   # unaligned movie: the older strip kernels
            if self._tc_host is None and self.strips_tc is None and self._regular is not None:
                rows, cols = self._regular
                self._tc_host = ops.make_strips_tc(rows, cols, self.bh, self.bw, self.d1, self.d2, self.ranks_host, self.col0_host,
                                                   self.bg.shape[0])
                if self._tc_host is None and self.strips is None:
                    self._build_simt_strips()
        self._finish_tc()
        if self.strips_tc is not None:
            if ops.project_stream_tc_ok(movie2d, self.d2, mean, inv_std):
                ops.project_stream_tc(movie2d, self.d2, self.strips_tc, self.bimg, mean, inv_std, z[: self.n_local], z[self.n_local :])
                _submark("projection.stream")
                return
            if self.strips is None and self._regular is not None:
                self._build_simt_strips()
        if self.strips is not None:
            ops.project_stream(movie2d, self.d2, self.strips, self.upack, mean, inv_std, z[: self.n_local], z[self.n_local :])
            _submark("projection.stream")
            return
        if self.n_local > 0:
            if self.supertiles is not None:
                ops.project_supertile(movie2d, self.d2, self.supertiles, self.bh, self.bw, self.uvals32, mean, inv_std,
                                      z[: self.n_local])
            else:
                if self.bh * self.bw > 512:
                    z[: self.n_local, :n].zero_()
                ops.project_local(movie2d, self.d2, self.starts_dev, self.bh, self.bw, self.ranks_dev, self.col0_dev,
                                  self.tasks, self.uvals32, mean, inv_std, z[: self.n_local])
            _submark("projection.local")
        zb = z[self.n_local :]
        zb[:, :n].zero_()
        ops.project_dense(movie2d, self.bg, mean, inv_std, zb)


def compute_lowrank_factorized_svd(u, v, only_left=False, factor="eigh", before_wait=None):
    """decomposition.py:936-1010 on the GPU.  `u` is a SparseU (or a scipy sparse matrix, converted),
    `v` a dense (R, t') tensor/array.  Returns the spatial mixing matrix P (R, k) (device tensor) such
    that U P has orthonormal columns; with only_left=False also (s, Vt) of the factorised product.

    factor="eigh" reproduces the reference's P = M E diag(1/sqrt(lambda)) (columns ordered by eigenvalue).
    factor="chol" (used by localmd_decomposition, whose next step re-diagonalises anyway) takes P = M L^-T
    from the float64 Cholesky factor G = L L^T: the same column space, hence the same final (R, s, Vt) in exact
    arithmetic, without the t' x t' eigenproblem; it falls back to "eigh" when G is numerically singular."""
    if not isinstance(u, SparseU):
        u = sparse_u_from_scipy(u)
    dev = u.uvals32.device
    v = _as_dev(v, dev)
    R = u.n_cols
    right = v.to(torch.float64) if R > v.shape[1] else torch.eye(R, dtype=torch.float64, device=dev)
    z = u.utu_times_f64(right)  # (R, m) float64
    _submark("whiten.utu")
    # G = M^T (U^T U M): own FP64 tensor-core kernel, upper tiles only, mirrored (exactly symmetric)
    g = ops.sym_product_f64(right, z, layout=1)
    _submark("whiten.gram")
    mix64 = None
    if factor == "chol" and only_left:
        chol, info = torch.linalg.cholesky_ex(g)
        dg = torch.diagonal(chol)
        ok = (info == 0) & (dg.min() > 1e-6 * dg.max())  # cond(G) < ~1e12 : safe in float64
        # the triangular solve is enqueued before the verdict is read back (the read is a host synchronisation; the device
        # keeps working through it) and discarded in the rare singular case
        spec = torch.linalg.solve_triangular(chol.t(), right, upper=True, left=False)   # X L^T = M (2.0 ms; L X^T = M^T: 2.7 ms)
        if before_wait is not None:     # the launching thread is about to wait for the device: the caller's cue for host work
            before_wait()
        if bool(ok.item()):
            mix64 = spec
            _submark("whiten.chol")
        del spec
    if mix64 is None:
        vals, vecs = sym_eigh_desc_abs(g)
        _submark("whiten.eigh")
        # "eig_vals > 0" (decomposition.py:988) evaluated in float64: drop what is numerically zero
        good = vals > vals[0] * 1e-13
        vals, vecs = vals[good], vecs[:, good]
        mix64 = torch.matmul(right, vecs) / torch.sqrt(vals)[None, :]
    mix = mix64.to(torch.float32)
    if only_left:
        return mix
    new_temporal = torch.matmul(mix64.t(), u.utu_times_f64(v.to(torch.float64))).to(torch.float32)
    return projected_svd(mix, new_temporal)


def sparse_u_from_scipy(u, device=None):
    """Wrap an arbitrary scipy sparse (d, R) matrix as a SparseU made of dense columns only.  Meant for
    the re-exported helper API on small problems, not for the main path."""
    device = torch.device(device if device is not None else "cuda")
    dense = torch.from_numpy(np.asarray(scipy.sparse.csr_matrix(u).todense(), dtype=np.float32)).to(device)
    d, R = dense.shape
    empty = np.zeros(0, dtype=np.int32)
    return SparseU(
        np.zeros((0, 2), np.int32), torch.zeros((0, 2), dtype=torch.int32, device=device), 2, 2, d, 1, empty.astype(np.int64),
        torch.zeros(0, dtype=torch.int32, device=device), torch.zeros((0, 4), dtype=torch.float64, device=device),
        torch.zeros((0, 4), dtype=torch.float32, device=device), dense.t().contiguous(),
    )


def projected_svd(projection, data, group=None, after_gram=None):
    """decomposition.py:1013-1137: Gram-based SVD of `data` (k, n) and R = projection @ left.
    With `group`, `data` holds this rank's frame columns: the k x k Gram is all-reduced and the
    returned Vt is the local column block.  after_gram: optional callable invoked once the Gram has been enqueued (the
    driver starts side-stream work there that should overlap the latency-bound eigensolver, not the GEMM)."""
    dev = data.device
    projection = _as_dev(projection, dev)
    k, n = data.shape
    n_total = n
    if group is not None:
        import torch.distributed as dist

        nt = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(nt, group=group)
        n_total = int(nt.item())
    if k <= n_total:
        # float64 Gram of the float32 rows on the FP64 tensor cores (own kernel: operands converted while they are staged,
        # tiles on and above the diagonal only, both triangles written from them -> exactly symmetric)
        gram = ops.sym_product_f64(data if data.stride(1) == 1 else data.contiguous())
        if group is not None:
            dist.all_reduce(gram, group=group)
        _submark("final_svd.gram")
        if after_gram is not None:
            after_gram()
        vals, left = sym_eigh_desc_abs(gram)
        _submark("final_svd.eigh")
        sing = torch.sqrt(vals).to(torch.float32)
        left = left.to(torch.float32)
        div = torch.where(sing == 0, torch.ones_like(sing), sing)
        # the two large float32 GEMMs of the stage run on the tensor cores at float32 accuracy (3xTF32)
        right = ops.matmul_3xtf32_any(left.t(), data) / div[:, None]
        return ops.matmul_3xtf32_any(projection, left), sing, right
    if group is not None:
        # k > T (fewer frames than components: tiny movies).  The T x T Gram couples every pair of frames, so the local column
        # blocks are exchanged (k x T floats, small by construction), every rank solves the same problem and keeps its own
        # block of Vt.
        rank, world = sharding.dist_info(group)
        mine = torch.tensor([n], dtype=torch.int64, device=dev)
        alln = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(alln, mine, group=group)
        counts = [int(x.item()) for x in alln]
        full = sharding.ragged_all_gather(data.t().contiguous(), counts, group).t().contiguous()   # (k, T)
        rmix, sing, vt = projected_svd(projection, full)
        off = sum(counts[:rank])
        return rmix, sing, vt[:, off : off + n].contiguous()
    d64 = data.to(torch.float64)
    gram = torch.matmul(d64.t(), d64)
    gram = 0.5 * (gram + gram.t())
    vals, right_t = sym_eigh_desc_abs(gram)
    sing = torch.sqrt(vals).to(torch.float32)
    right_t = right_t.to(torch.float32)
    div = torch.where(sing == 0, torch.ones_like(sing), sing)
    left = torch.matmul(data, right_t / div[None, :])
    return torch.matmul(projection, left), sing, right_t.t().contiguous()


# ---------------------------------------------------------------------------------------------
# driver
# ---------------------------------------------------------------------------------------------
def localmd_decomposition(
    dataset_obj,
    block_sizes=None,
    frame_range=None,
    max_components: int = 50,
    background_rank: int = 15,
    sim_conf: int = 5,
    frame_batch_size: int = 10000,
    dtype: str = "float32",
    num_workers: int = 0,
    pixel_batch_size: int = 5000,
    max_consecutive_failures=1,
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
    *,
    block_height: Optional[int] = None,
    block_width: Optional[int] = None,
    frames_to_init: Optional[int] = None,
    draws=None,
    seed: Optional[int] = None,
    device=None,
    timings: Optional[dict] = None,
    details: Optional[dict] = None,
    verbose: bool = False,
    group=None,
):
    if block_sizes is None:
        if block_height is None or block_width is None:
            raise TypeError("block_sizes (or block_height and block_width) is required")
        block_sizes = [block_height, block_width]
    if frame_range is None:
        if frames_to_init is None:
            raise TypeError("frame_range (or frames_to_init) is required")
        frame_range = frames_to_init
    if dtype != "float32":
        raise ValueError("only dtype='float32' is supported (the reference computes in float32 on device)")
    for name_, fn_ in (("spatial_denoiser", spatial_denoiser), ("temporal_denoiser", temporal_denoiser)):
        if fn_ is not None and not callable(fn_):
            raise TypeError("%s must be a callable on CUDA torch tensors" % name_)
    if not torch.cuda.is_available():
        raise RuntimeError("localmd_b200 needs a CUDA device (built for sm_100a); there is no CPU fallback")
    dev = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
    say = display if verbose else (lambda m: None)

    T, d1, d2 = (int(x) for x in dataset_obj.shape)
    d = d1 * d2
    check_fov_size((d1, d2))
    if group is not None and seed is None:
        # every rank must draw the same frame lists and sketches (they size the collectives below): rank 0 picks the seed
        import torch.distributed as dist

        seed_t = torch.tensor([int(np.random.SeedSequence().entropy % (2**62))], dtype=torch.int64, device=dev)
        dist.broadcast(seed_t, src=dist.get_global_rank(group, 0), group=group)
        seed = int(seed_t.item())
    rng = np.random.default_rng(seed)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(rng.integers(0, 2**62)))
    draws = draws if draws is not None else object()
    take = lambda name: getattr(draws, name, None)  # noqa: E731

    with torch.cuda.device(dev), ops.fp32_matmul(), _short_switch_interval():
        global _ACTIVE_TIMER
        tm = _Timer(timings, dev)
        _ACTIVE_TIMER = tm
        tm.mark("start")
        # ---- frame sharding (group = torch.distributed process group, one rank per GPU) ------------
        rank, world = sharding.dist_info(group)
        if isinstance(dataset_obj, DeviceMovie):
            movie = dataset_obj
        elif group is None:
            movie = DeviceMovie(dataset_obj, dev, batch_frames=max(1024, frame_batch_size))
        else:
            lo, hi = sharding.shard_bounds(T, world)[rank]
            movie = DeviceMovie(dataset_obj, dev, batch_frames=max(1024, frame_batch_size), frame_lo=lo, frame_hi=hi)
        bounds = None
        if group is not None:
            import torch.distributed as dist

            mine = torch.tensor([movie.lo, movie.hi], dtype=torch.int64, device=dev)
            allb = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allb, mine, group=group)
            bounds = [(int(b[0]), int(b[1])) for b in torch.stack(allb).cpu()]
        tm.mark("upload")

        # ---- thresholds (decomposition.py:706-711), enqueued early on a high-priority side stream ------------------
        # The simulation only depends on the block geometry: its ~4 ms of small launches run beside the statistics pass.
        wc_eff = frame_range if window_chunks is None else window_chunks
        fr_eff = min(frame_range, T)
        wc_eff = fr_eff if fr_eff <= wc_eff else wc_eff
        bh_e, bw_e = (min(int(b_), int(n_)) for b_, n_ in zip(block_sizes, (d1, d2)))
        thr_pending = None
        if take("thresholds") is None and min(bh_e, bw_e) >= 10:
            side_thr = _side_stream(dev, -1)
            side_thr.wait_stream(torch.cuda.current_stream(dev))
            gen_sim = torch.Generator(device=dev)
            gen_sim.manual_seed(gen.initial_seed() + 1)
            with torch.cuda.stream(side_thr):
                thr_pending = (simulate_thresholds(bh_e, bw_e, wc_eff, sim_conf, draws, gen_sim, dev, defer=True), bh_e, bw_e, wc_eff)

        # ---- PMDLoader.__init__ : normalisers + background (pmd_loader.py:172-173) --------------
        say("Computing Video Statistics")
        mean, std = compute_mean_and_noise(movie, compute_normalizer, group)
        inv_std = 1.0 / std
        tm.mark("stats")
        if background_rank > 0:
            n_bg = min(1000, T)
            bg_frames = take("bg_frames")
            if bg_frames is None:
                bg_frames = rng.choice(T, size=n_bg, replace=False).tolist()
            bg_sketch = take("bg_sketch")
            bg_sketch = (
                _as_dev(bg_sketch, dev)
                if bg_sketch is not None
                else torch.randn((len(bg_frames), background_rank + 10), generator=gen, device=dev, dtype=torch.float32)
            )
            bg = background_basis(movie, mean, std, bg_frames, bg_sketch, background_rank, group, bounds)
        else:
            bg = torch.zeros((1, d), dtype=torch.float32, device=dev)
        tm.mark("background")

        # ---- frame selection (decomposition.py:678-693) ------------------------------------------
        if window_chunks is None:
            window_chunks = frame_range
        if T < frame_range:
            say("WARNING: Specified using more frames than there are in the dataset.")
            frame_range = T
            frames = list(range(T))
            if frame_range <= window_chunks:
                window_chunks = frame_range
        else:
            if frame_range <= window_chunks:
                window_chunks = frame_range
            init_frames = take("init_frames")
            if init_frames is not None:
                frames = [int(f) for f in init_frames]
            else:
                frames = identify_window_chunks(frame_range, T, window_chunks, rng)
        say("We are initializing on a total of {} frames".format(len(frames)))

        block_sizes = update_block_sizes(block_sizes, (d1, d2))
        bh, bw = block_sizes

        # ---- thresholds (decomposition.py:706-711) ------------------------------------------------
        thr = take("thresholds")
        if thr is not None:
            thr_s, thr_t = float(thr[0]), float(thr[1])
        elif thr_pending is not None and thr_pending[1:] == (bh, bw, window_chunks):
            thr_s = thr_t = thr_pending[0]   # callable, resolved at the rank decision of the block stage
        else:
            thr_s, thr_t = simulate_thresholds(bh, bw, window_chunks, sim_conf, draws, gen, dev)
        tm.mark("thresholds")

        # ---- init frames: standardise + background removal (pmd_loader.py:348-389) ----------------
        # Only the first crop = (t // taf) * taf init frames are ever used (decomposition.py:773-774); they
        # are written straight into the pixel-major layout of the block kernels.
        t_init = len(frames)
        if temporal_avg_factor >= t_init:
            raise ValueError("Need at least {} frames".format(temporal_avg_factor))
        if t_init // temporal_avg_factor <= max_components:
            say("WARNING: temporal avg factor is too big, max rank per block adjusted to {}.".format(t_init // temporal_avg_factor))
            max_components = int(t_init // temporal_avg_factor)
        crop = (t_init // temporal_avg_factor) * temporal_avg_factor
        r = int(max_components)
        if r + 10 > 112:
            raise ValueError("max_components > 102 is not supported by the sm_100a Jacobi kernel")
        dim_1_iters, dim_2_iters = tile_starts(d1, bh), tile_starts(d2, bw)
        starts = np.stack(np.meshgrid(dim_1_iters, dim_2_iters, indexing="ij"), axis=-1).reshape(-1, 2).astype(np.int32)
        nb = starts.shape[0]
        # summed pyramid weights of the covering blocks: sum_b shift(block_weights) = Rm^T W Cm with the 0/1
        # incidence matrices of (block-local row, FOV row) and (block-local column, FOV column); exact in float64.
        # (Rank independent host work: done HERE, before the init filter is enqueued -- the table uploads further down drain
        # the stream, and 1.2 ms of NumPy between that drain and the first block kernel left the device idle.)
        block_weights = pyramid_weights(bh, bw)
        rm, cm = np.zeros((bh, d1)), np.zeros((bw, d2))
        for q in range(bh):
            rm[q, np.asarray(dim_1_iters) + q] = 1.0
        for q in range(bw):
            cm[q, np.asarray(dim_2_iters) + q] = 1.0
        cumw = rm.T @ block_weights.astype(np.float64) @ cm
        # blocks are partitioned over the ranks (contiguous index ranges); results are all-gathered below
        b0, b1 = sharding.block_partition(nb, world)[rank]
        row_lo, row_hi = 0, d1                                   # image rows of the init movie this rank holds
        if group is None:
            src2d, idx = movie.frame_source(frames[:crop])
            yt = ops.standardize_frames_t(src2d, idx, mean, std)  # (d, ld)
            del src2d, idx
            own_lo, own_hi = 0, d1
        else:
            # Patch-partitioned init window: this rank receives from the window's owner rank(s) only the
            # image rows its blocks cover and standardises those; sums over the field of view (the background traces) are
            # taken over a disjoint row cover and all-reduced.
            parts = sharding.block_partition(nb, world)
            ranges = sharding.block_row_ranges(starts, bh, parts)
            if any(hi_ <= lo_ for lo_, hi_ in ranges):            # more ranks than blocks: every rank takes the whole window
                ranges = [(0, d1)] * world
                owned = [(0, d1) if r_ == 0 else (0, 0) for r_ in range(world)]
            else:
                owned = sharding.owned_row_ranges(ranges, d1)
            row_lo, row_hi = ranges[rank]
            own_lo, own_hi = owned[rank]
            src2d = sharding.exchange_frame_rows(movie, frames[:crop], ranges, d2, group, bounds)
            idx = torch.arange(src2d.shape[0], dtype=torch.int64, device=dev)
            yt = ops.standardize_frames_t(src2d, idx, mean[row_lo * d2 : row_hi * d2].contiguous(),
                                          std[row_lo * d2 : row_hi * d2].contiguous())  # (local pixels, ld)
            del src2d, idx
        bg_loc = bg[:, row_lo * d2 : row_hi * d2].contiguous() if (row_lo, row_hi) != (0, d1) else bg
        if group is not None:
            import torch.distributed as dist

            o0, o1 = (own_lo - row_lo) * d2, (own_hi - row_lo) * d2
            if o1 > o0:
                bo = bg_loc[:, o0:o1].contiguous()
                vbg = ops.bg_project_t(yt[o0:o1], bo) if bo.shape[0] <= 16 else torch.matmul(bo, yt[o0:o1]).contiguous()
            else:
                vbg = torch.zeros((bg.shape[0], yt.shape[1]), dtype=torch.float32, device=dev)
            dist.all_reduce(vbg, group=group)
            if bg.shape[0] <= 16:
                ops.bg_remove_t(yt, bg_loc, vbg)
            else:
                yt.addmm_(bg_loc.t(), vbg, alpha=-1.0)
        elif bg.shape[0] <= 16:   # skinny contractions: one streaming pass over yt each (csrc/bgfilter.cu)
            vbg = ops.bg_project_t(yt, bg)  # (K, ld)
            ops.bg_remove_t(yt, bg, vbg)
        else:
            vbg = torch.matmul(bg, yt).contiguous()  # (K, ld)
            yt.addmm_(bg.t(), vbg, alpha=-1.0)
        if pixel_weighting is not None:
            pw = _as_dev(np.asarray(pixel_weighting, dtype=np.float32).reshape(-1), dev)
            yt *= pw[row_lo * d2 : row_hi * d2][:, None]
        tm.mark("init_filter")

        starts_dev = ops.h2d(starts, dev)
        # block origins relative to the rows of the init movie this rank holds
        starts_fit = starts_dev if row_lo == 0 else (starts_dev - ops.h2d(np.array([row_lo, 0], dtype=np.int32), dev))
        block_weights_dev = ops.h2d(block_weights.reshape(-1), dev)
        cumw_dev = ops.h2d(cumw.reshape(-1), dev)

        # ---- block fits (decomposition.py:790-838) -------------------------------------------------
        bs = take("block_sketches")
        windowed = window_chunks < crop
        if not windowed:
            if bs is not None:
                sketches = torch.stack([_as_dev(b_[0] if isinstance(b_, (list, tuple)) else b_, dev) for b_ in bs])
            else:
                sketches = torch.randn((nb, crop // temporal_avg_factor, r + 10), generator=gen, device=dev, dtype=torch.float32)
            u_blk, v_blk, ranks_loc, sstat, tstat = block_decompositions(
                yt, crop, d2, starts_fit[b0:b1].contiguous(), bh, bw, r, temporal_avg_factor, spatial_avg_factor, thr_s, thr_t,
                int(max_consecutive_failures), sketches[b0:b1], spatial_denoiser, temporal_denoiser,
            )
        else:
            # one Gaussian per (block, visited window) (decomposition.py:475, 62)
            n_win = len(range(0, crop, window_chunks))
            if bs is not None:
                # a block that reached max_components stops consuming sketches: missing entries are never used
                shape_w = (window_chunks // temporal_avg_factor, r + 10)
                sketches = [torch.stack([_as_dev(b_[wi], dev) if wi < len(b_) else torch.zeros(shape_w, dtype=torch.float32, device=dev)
                                         for b_ in bs])[b0:b1] for wi in range(n_win)]
            else:
                sketches = [torch.randn((nb, window_chunks // temporal_avg_factor, r + 10), generator=gen, device=dev,
                                        dtype=torch.float32)[b0:b1] for _ in range(n_win)]
            u_blk, v_blk, ranks_loc, sstat, tstat = block_decompositions_windowed(
                yt, crop, d2, starts_fit[b0:b1].contiguous(), bh, bw, r, temporal_avg_factor, spatial_avg_factor, thr_s, thr_t,
                int(max_consecutive_failures), sketches, int(window_chunks), spatial_denoiser, temporal_denoiser,
            )
        del sketches
        if group is None:
            ranks_dev = ranks_loc
        else:
            bcounts = [hi_ - lo_ for lo_, hi_ in sharding.block_partition(nb, world)]
            ranks_dev = sharding.ragged_all_gather(ranks_loc.contiguous(), bcounts, group)
            if details is not None:
                sstat = sharding.ragged_all_gather(sstat.contiguous(), bcounts, group)
                tstat = sharding.ragged_all_gather(tstat.contiguous(), bcounts, group)
        ranks_host = ranks_dev.cpu().numpy().astype(np.int64)
        _ht("ranks_host")
        tm.mark("blocks")

        # ---- weighted sparse assembly (decomposition.py:811-857) -----------------------------------
        # this rank's kept components: weighted values (float64 + float32) and temporal traces
        ranks_loc_host = ranks_host[b0:b1]
        col0_loc = ops.h2d(np.concatenate([[0], np.cumsum(ranks_loc_host)[:-1]]).astype(np.int64), dev)
        ncol_loc = int(ranks_loc_host.sum())
        uv64, uv32 = ops.assemble_u(
            u_blk, bh, bw, starts_dev[b0:b1].contiguous(), ranks_loc, col0_loc, block_weights_dev, cumw_dev, d2, ncol_loc,
        )
        # (output_size: without it repeat_interleave reads the total back from the device -- a host synchronisation)
        blk_of_col = torch.repeat_interleave(torch.arange(b1 - b0, device=dev), ranks_loc.to(torch.int64), output_size=ncol_loc)
        _ht("assemble_u enqueued")
        comp_of_col = torch.arange(ncol_loc, device=dev) - col0_loc[blk_of_col]
        v_loc = v_blk[blk_of_col, comp_of_col][:, :crop].contiguous()  # (local columns, t)
        del u_blk, v_blk, yt
        if group is not None:
            ccounts = [int(ranks_host[lo_:hi_].sum()) for lo_, hi_ in sharding.block_partition(nb, world)]
            uv64 = sharding.ragged_all_gather(uv64, ccounts, group)
            uv32 = sharding.ragged_all_gather(uv32, ccounts, group)
            v_loc = sharding.ragged_all_gather(v_loc, ccounts, group)
        su = SparseU(starts, starts_dev, bh, bw, d1, d2, ranks_host, ranks_dev, uv64, uv32, bg)
        _ht("SparseU built")
        v_init = torch.cat([v_loc, vbg[:, :crop]], dim=0)  # (R, t)
        del v_loc
        say("The total rank before pruning is {}".format(su.n_cols))
        if timings is not None:
            timings["__info__"] = dict(n_cols=int(su.n_cols), n_local=int(su.n_local), nb=int(nb), mean_rank=float(ranks_host.mean()),
                                       max_rank=int(ranks_host.max()))
        tm.mark("assemble")
        _ht("assemble mark")

        # ---- result CSR (decomposition.py:811-857, 912-933) --------------------------------------------
        # Written straight from the block-component form by two small kernels (count, fill) as soon as the components are
        # assembled; the worst-case buffers are sliced to the true entry count at the very end (finish_export).
        if order == "F":     # row of U of physical pixel (i, j): i + j * d1 (pmdarray.py:37-38), formed on the device
            row_ids = (torch.arange(d1, device=dev)[:, None] + d1 * torch.arange(d2, device=dev)[None, :]).reshape(-1).contiguous()
        else:
            row_ids = torch.arange(d, device=dev)
        exported = su.export_csr(row_ids)

        # ---- orthogonalisation (decomposition.py:860-881) -------------------------------------------
        if rank_prune:
            if rank_prune_factor <= 0 or rank_prune_factor > 1:
                raise ValueError("Rank prune factor should be a value in the interval (0, 1]")
            shape = (v_init.shape[1], int(min(su.n_cols, v_init.shape[1]) * rank_prune_factor))
            ps = take("prune_sketch")
            if ps is not None:
                ps = ps(shape) if callable(ps) else ps
                ps = _as_dev(ps, dev)
                if tuple(ps.shape) != shape:
                    raise ValueError("prune_sketch has shape %s, expected %s" % (tuple(ps.shape), shape))
            else:
                ps = torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)
            # the prune sketch V Omega (R x t)(t x k'): float32-accurate on the tensor cores (3xTF32) instead of a SIMT GEMM
            _ht("prune sketch drawn")
            sketch = ops.matmul_3xtf32_any(v_init, ps)
            _ht("prune GEMM enqueued")
            su.start_host_tables()   # worker thread: U^T U bookkeeping + K7 strip tables, beside the GEMMs just enqueued
            p = compute_lowrank_factorized_svd(su, sketch, only_left=True, factor="chol")
            del sketch
        else:
            p = compute_lowrank_factorized_svd(su, v_init, only_left=True, factor="chol")
        say("After performing rank reduction, the updated rank is {}".format(p.shape[1]))
        tm.mark("whiten")
        _ht("whiten enqueued")

        # ---- full-movie projection (pmd_loader.py:316-346) ------------------------------------------
        v_full = project_movie(movie, su, p, mean, inv_std)
        tm.mark("projection")
        _ht("projection enqueued")

        # ---- final SVD (decomposition.py:896-904) ---------------------------------------------------
        rmix, s, vt = projected_svd(p, v_full, group)
        good = s != 0
        rmix, s, vt = rmix[:, good], s[good], vt[good, :]
        if group is not None:  # Vt column shards -> the full (k, T) factor on every rank
            fcounts = [hi_ - lo_ for lo_, hi_ in bounds]
            vt = sharding.ragged_all_gather(vt.t().contiguous(), fcounts, group).t().contiguous()
        tm.mark("final_svd")
        _ht("final_svd enqueued")

        # ---- result object ---------------------------------------------------------------------------
        (indptr, indices, values), csr32 = SparseU.finish_export(exported)
        out = PMDArray._from_device((indptr, indices, values), csr32, rmix.contiguous(), s.contiguous(),
                                    vt.contiguous(), (T, d1, d2), order, mean, std, dev)
        tm.mark("export")
        tm.finish()
        if _HOST_TRACE is not None:
            t0_ = _HOST_TRACE[0][1]
            sys.stderr.write("host trace (ms): " + ", ".join("%s %.2f" % (l_, 1e3 * (t_ - t0_)) for l_, t_ in _HOST_TRACE) + "\n")
            _HOST_TRACE.clear()
        if timings is not None and "__info__" in timings:
            timings["__info__"]["h2d_bytes"] = int(movie.h2d_bytes)
        _ACTIVE_TIMER = None
        if details is not None:
            details.update(
                ranks=ranks_host.astype(np.int32), block_starts=[tuple(x) for x in starts.tolist()], thresholds=thr_s() if callable(thr_s) else (thr_s, thr_t),
                spatial_basis=np.stack([img.reshape(-1, order=order) for img in bg.cpu().numpy().reshape(-1, d1, d2)], axis=1),
                sstat=sstat.cpu().numpy(), tstat=tstat.cpu().numpy(), mixing=p.cpu().numpy(), v_init=v_init.cpu().numpy(),
                v_full=v_full.cpu().numpy(), frames=frames, h2d_bytes=movie.h2d_bytes,
            )
        return out


def project_movie(movie: DeviceMovie, su: SparseU, p, mean, inv_std):
    """K7: V = P^T U^T ((Y - mean)/std) over this rank's frames -> (k, n_local) float32."""
    dev = movie.device
    k = p.shape[1]
    v_full = torch.empty((k, movie.n_local), dtype=torch.float32, device=dev)
    # the contraction dimension (columns of U) is padded to a multiple of 8 so that both GEMM operands have
    # 16-byte aligned rows (tensor-core kernels need it); the padding rows / columns are zero
    rpad = (su.n_cols + 7) // 8 * 8
    pt = torch.zeros((k, rpad), dtype=torch.float32, device=dev)
    pt[:, : su.n_cols] = p.t()
    # mixing GEMM P^T Z at float32 accuracy as ONE TF32 + ONE bf16 library GEMM (ops.matmul_split): operands split once
    pt_pair = ops.split_pairs(pt, 1)
    for f0, chunk in movie.batches():
        n = chunk.shape[0]
        npad = (n + 7) // 8 * 8
        z = torch.empty((rpad, npad), dtype=torch.float32, device=dev)
        z[su.n_cols :].zero_()
        if npad != n:
            z[:, n:].zero_()
        su.project(chunk, mean, inv_std, z[: su.n_cols])
        _submark("projection.dense")
        z_pair = ops.split_pairs(z, 0)
        v_full[:, f0 : f0 + n].copy_(ops.matmul_split(pt, pt_pair, z, z_pair)[:, :n])
        del z_pair
        _submark("projection.mix")
        del z
    return v_full
