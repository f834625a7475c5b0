"""Generate the golden fixtures by executing the UNMODIFIED reference (/root/reference/localmd)
over the NumPy-backed JAX stand-in in tests/golden/jax_shim.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs tests/golden/case_<name>.npz.  Nothing at test/bench time reads /root/reference.

The reference source is imported as-is; the only instrumentation is *recording* (wrappers around
np.random.choice, PMDLoader and windowed_pmd that call the originals and keep their results).
Each fixture stores: the parameters, how to rebuild the input movie (tests/synth.make_movie args +
a checksum), every random draw in consumption order ((seed, shape) for Gaussians -- regenerate with
oracle.pmd_oracle.normal_from_seed -- and the literal np.random.choice results), and the outputs
(U as CSR arrays, R, s, Vt, mean_img, noise_var_img, per-block ranks, thresholds, background basis).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jax_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(HERE))

import jax  # noqa: E402  (the shim)
import localmd.decomposition as ref_dec  # noqa: E402
import localmd.pmd_loader as ref_loader  # noqa: E402
from synth import make_movie  # noqa: E402

CASES = {
    # multi-chunk stats (1024 + 276 -> second chunk has exactly one Welch segment), F order
    "main_F": dict(
        movie=dict(T=1300, d1=40, d2=36, n_cells=3, seed=11, dtype="float32"),
        block_sizes=[16, 16],
        frame_range=600,
        kwargs=dict(max_components=12, background_rank=2, frame_batch_size=256, pixel_batch_size=400),
    ),
    # uint16 input, C order, non-square blocks, rank pruning, short tail chunk (mean only)
    "prune_C_u16": dict(
        movie=dict(T=1100, d1=30, d2=50, n_cells=12, seed=12, dtype="uint16"),
        block_sizes=[12, 20],
        frame_range=200,
        kwargs=dict(
            max_components=6,
            background_rank=1,
            frame_batch_size=300,
            pixel_batch_size=500,
            order="C",
            rank_prune=True,
            rank_prune_factor=0.5,
            temporal_avg_factor=5,
        ),
    ),
    # fewer than 256 frames (no normaliser), frame_range > T, block larger than FOV, no background
    "tiny_noNorm": dict(
        movie=dict(T=200, d1=24, d2=20, n_cells=3, seed=13, dtype="float32"),
        block_sizes=[32, 32],
        frame_range=5000,
        kwargs=dict(max_components=5, background_rank=0, frame_batch_size=64, pixel_batch_size=100),
    ),
    # R > t: the whitening uses right_mat = V (decomposition.py:976-977)
    "wide_R": dict(
        movie=dict(T=600, d1=40, d2=36, n_cells=40, seed=14, dtype="float32", noise=0.3),
        block_sizes=[16, 16],
        frame_range=100,
        kwargs=dict(
            max_components=8, background_rank=2, frame_batch_size=200, pixel_batch_size=400, temporal_avg_factor=5,
            max_consecutive_failures=2,
        ),
    ),
    # window_chunks < frame_range: residual windows (decomposition.py:333-387, 455-515)
    "windows": dict(
        movie=dict(T=1300, d1=32, d2=32, n_cells=8, seed=15, dtype="float32"),
        block_sizes=[16, 16],
        frame_range=400,
        kwargs=dict(max_components=6, background_rank=1, frame_batch_size=500, pixel_batch_size=400, window_chunks=200),
    ),
}


def run_case(name, spec):
    mv = dict(spec["movie"])
    dtype = np.dtype(mv.pop("dtype"))
    movie = make_movie(dtype=dtype, **mv)
    np.random.seed(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
    jax.random.RANDOM_LOG.clear()

    rec = dict(choice=[], loader=None, blocks=[])
    orig_choice = np.random.choice

    def choice(*a, **k):
        out = orig_choice(*a, **k)
        rec["choice"].append(np.array(out))
        return out

    orig_loader = ref_dec.PMDLoader

    class RecLoader(orig_loader):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            rec["loader"] = dict(spatial_basis=np.array(self.spatial_basis))

    orig_wp = ref_dec.windowed_pmd

    def wp(*a, **k):
        n_before = len(jax.random.RANDOM_LOG)
        sc, tc = orig_wp(*a, **k)
        rec["blocks"].append(dict(rank=sc.shape[2], n_sketch=len(jax.random.RANDOM_LOG) - n_before, u=sc, v=tc))
        return sc, tc

    orig_thr = ref_dec.threshold_heuristic

    def thr(*a, **k):
        out = orig_thr(*a, **k)
        rec["thresholds"] = (float(out[0]), float(out[1]))
        rec["n_log_after_thr"] = len(jax.random.RANDOM_LOG)
        return out

    np.random.choice = choice
    ref_dec.PMDLoader = RecLoader
    ref_dec.windowed_pmd = wp
    ref_dec.threshold_heuristic = thr
    try:
        arr = ref_dec.localmd_decomposition(movie, spec["block_sizes"], spec["frame_range"], **spec["kwargs"])
    finally:
        np.random.choice = orig_choice
        ref_dec.PMDLoader = orig_loader
        ref_dec.windowed_pmd = orig_wp
        ref_dec.threshold_heuristic = orig_thr

    log = list(jax.random.RANDOM_LOG)
    seeds = np.array([s for s, _ in log], dtype=np.int64)
    shapes = np.full((len(log), 3), -1, dtype=np.int64)
    for i, (_, shp) in enumerate(log):
        shapes[i, : len(shp)] = shp
    u = arr.u
    assert u.has_canonical_format or True
    u.sort_indices()
    sample_frames = np.array([0, movie.shape[0] // 2, movie.shape[0] - 1])
    recon = arr[sample_frames.tolist(), :, :]
    out = dict(
        spec=json.dumps(spec),
        movie_checksum=np.array([float(movie.astype(np.float64).sum()), float(movie[3, 5, 7])]),
        normal_seeds=seeds,
        normal_shapes=shapes,
        n_log_after_thr=np.array(rec.get("n_log_after_thr", -1)),
        n_choice=np.array(len(rec["choice"])),
        thresholds=np.array(rec["thresholds"]),
        spatial_basis=rec["loader"]["spatial_basis"],
        block_ranks=np.array([b["rank"] for b in rec["blocks"]], dtype=np.int32),
        block_n_sketch=np.array([b["n_sketch"] for b in rec["blocks"]], dtype=np.int32),
        block0_u=rec["blocks"][0]["u"].astype(np.float32),
        block0_v=rec["blocks"][0]["v"].astype(np.float32),
        U_data=u.data,
        U_indices=u.indices,
        U_indptr=u.indptr,
        U_shape=np.array(u.shape),
        R=np.asarray(arr.r),
        s=np.asarray(arr.s),
        Vt=np.asarray(arr.v),
        mean_img=np.asarray(arr.mean_img),
        noise_var_img=np.asarray(arr.var_img),
        fov_order=np.array(arr.order),
        recon_frames=sample_frames,
        recon=recon,
        pixel_trace=arr[:, 5, 7],
        crop=arr[10:20, 3:9, 4:15],
    )
    for i, c in enumerate(rec["choice"]):
        out["choice_%d" % i] = c
    path = os.path.join(HERE, "case_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(
        name, "-> U", u.shape, "nnz", u.nnz, "R", arr.r.shape, "Vt", arr.v.shape, "ranks", out["block_ranks"].tolist(),
        "thr", rec["thresholds"], "size %.2f MB" % (os.path.getsize(path) / 1e6),
    )


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n, CASES[n])
