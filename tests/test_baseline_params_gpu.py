"""Parity at the PRODUCTION parameters of BASELINE.json's configurations: max_components = 50 (sketch width 60),
background rank 15, rank_prune 0.33 and 20x20 / 32x32 / 40x40 blocks (configs[1], [2], [3]) on fields of view large
enough that the tensor-core kernels of the real path run (project_ts_kernel, block_project_tc, block_spatial_tc,
the float32 Jacobi sweeps of n = 60, the Cholesky whitening of a few hundred columns).

The CUDA path (through the C ABI) and the CPU oracle get the same movie and the same host-supplied draws.  Tolerances
are the ones BASELINE.json's north_star states, checked against the reference algorithm evaluated in float64
(oracle.precision(np.float64)): ranks / CSR structure bit-exact except blocks with a statistic within EPS_STAT of a
threshold, singular values <= 1e-4 relative on the lead set, principal angles of U.R and Vt <= 1e-3 rad, Y_hat <= 1e-4
relative Frobenius error.  The float32 oracle is reported beside it (its own distance from float64 is the slack)."""
import numpy as np
import pytest

import oracle.pmd_oracle as O
from oracle import parity as P
from synth import make_movie

pytestmark = pytest.mark.gpu

# name: (T, d1, d2, block, t, n_cells, blob sigma)
CASES = {
    "c2_20x20": (2048, 128, 128, 20, 1000, 25, (3.0, 5.0)),
    "c3_32x32": (2048, 128, 160, 32, 1000, 25, (3.0, 5.0)),
    "c4_40x40": (1536, 160, 160, 40, 1000, 8, (10.0, 18.0)),
    # max_components = 80 (> the 64 components of one tensor-core accumulator tile; sketch width 90, Jacobi n = 90)
    "r80_32x32": (2048, 96, 128, 32, 1000, 20, (3.0, 5.0)),
}
RANK = {"r80_32x32": 80}
_cache = {}


def _inputs(name, seed=3):
    T, d1, d2, blk, t, n_cells, blob = CASES[name]
    movie = make_movie(T, d1, d2, n_cells=n_cells, seed=21, blob_sigma=blob)
    rng = np.random.default_rng(seed)
    nb = len(O.tile_starts(d1, blk)) * len(O.tile_starts(d2, blk))
    r, K = RANK.get(name, 50), 15
    prune_seed = int(rng.integers(0, 2**31))
    draws = O.Draws(
        bg_frames=rng.choice(T, min(1000, T), replace=False).tolist(),
        bg_sketch=rng.standard_normal((min(1000, T), K + 10), dtype=np.float32),
        init_frames=list(range(400, 400 + t)),
        # thresholds: oracle.threshold_heuristic on 250 noise blocks of these sizes (t = 1000); the simulation has its own tests
        thresholds={20: (1.3623216, 2.3522365), 32: (1.3852178, 2.348523), 40: (1.3879093, 2.3590899)}[blk],
        block_sketches=[[rng.standard_normal((t // 10, r + 10), dtype=np.float32)] for _ in range(nb)],
        prune_sketch=lambda shape: np.random.default_rng(prune_seed).standard_normal(shape, dtype=np.float32),
    )
    kw = dict(block_sizes=[blk, blk], frame_range=t, rank_prune=True, max_components=r, background_rank=K)
    return movie, draws, kw


def _run(name, pixel_weighting=None):
    key = (name, pixel_weighting is not None)
    if key not in _cache:
        import localmd_b200

        movie, draws, kw = _inputs(name)
        det = {}
        arr = localmd_b200.localmd_decomposition(movie, draws=draws, details=det, pixel_weighting=pixel_weighting, **kw)
        okw = dict(rank_prune=True, max_components=kw["max_components"], background_rank=kw["background_rank"],
                   pixel_weighting=pixel_weighting)
        ref32 = O.localmd_decomposition_oracle(movie, kw["block_sizes"], kw["frame_range"], draws, **okw)
        with O.precision(np.float64):
            ref64 = O.localmd_decomposition_oracle(movie, kw["block_sizes"], kw["frame_range"], draws, **okw)
        _cache[key] = (movie, arr, det, ref32, ref64)
    return _cache[key]


def movie_pixels(name):
    return CASES[name.split("+")[0]][1] * CASES[name.split("+")[0]][2]


def _check(name, arr, det, ref32, ref64):
    rep64 = P.parity_report(arr, det, ref64)
    rep32 = P.parity_report(arr, det, ref32)
    print(name, "vs float64 oracle:", rep64)
    print(name, "vs float32 oracle:", {k: rep32[k] for k in ("blocks_rank_equal", "blocks_rank_differ_within_eps",
                                                                 "blocks_rank_differ_outside_eps", "s_max_rel_err_lead",
                                                                 "yhat_rel_fro_err")})
    # ranks: bit-exact outside the stated band around a threshold, against BOTH evaluations of the reference algorithm
    assert rep64["blocks_rank_differ_outside_eps"] == 0, rep64
    assert rep32["blocks_rank_differ_outside_eps"] == 0, rep32
    assert det["ranks"].max() > 4, "the case is meant to exercise multi-slot blocks"
    assert arr.u.shape[0] == movie_pixels(name)
    if rep64["blocks_rank_differ_within_eps"] == 0:
        assert rep64["csr_indptr_equal"] and rep64["csr_indices_equal"]
        assert P.within_north_star(rep64), rep64
    return rep64


@pytest.mark.parametrize("name", sorted(CASES))
def test_production_parameters_match_oracle(name):
    movie, arr, det, ref32, ref64 = _run(name)
    rep = _check(name, arr, det, ref32, ref64)
    assert rep["blocks_total"] == len(det["ranks"])


def test_pixel_weighting_matches_oracle():
    """decomposition.py:717-718: the standardised, background-filtered init movie is multiplied by a per-pixel weight
    image before the block fits (the projection of the full movie is NOT weighted)."""
    name = "c2_20x20"
    T, d1, d2 = CASES[name][:3]
    yy, xx = np.mgrid[0:d1, 0:d2]
    w = (0.5 + 1.5 * np.exp(-((yy - d1 / 2) ** 2 + (xx - d2 / 3) ** 2) / (2 * 40.0**2))).astype(np.float32)
    movie, arr, det, ref32, ref64 = _run(name, pixel_weighting=w)
    _check(name + "+pixel_weighting", arr, det, ref32, ref64)
    # and the weighting does change the result (the test would pass vacuously if the argument were ignored)
    _, arr0, det0, _, ref0 = _run(name)
    assert not np.array_equal(ref64.ranks, ref0.ranks) or P.recon_rel_err(
        ref64.u, ref64.r, ref64.s, ref64.vt, ref0.u, ref0.r, ref0.s, ref0.vt) > 1e-4
