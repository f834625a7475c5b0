"""Known-answer tests the reference does not have: Welch estimator, tiling, pyramid
weights, rank rule, roughness statistics, pooling, frame batching, PMDArray slicing against the
reference's own pmdarray.py outputs stored in the fixtures."""
import numpy as np
import pytest

import oracle.pmd_oracle as O


def test_welch_white_noise_and_sinusoid():
    rng = np.random.default_rng(0)
    x = (3.0 * rng.standard_normal((64, 1024))).astype(np.float32)
    est = O.welch_noise_estimate(x)
    assert abs(est.mean() - 3.0) < 0.1
    t = np.arange(1024)
    sin = (5 * np.sin(2 * np.pi * 0.1 * t))[None].astype(np.float32)  # below fs/4
    assert O.welch_noise_estimate(sin)[0] < 1e-3
    np.testing.assert_allclose(O.welch_noise_estimate_explicit(x), est, rtol=2e-5)
    y = (200 + rng.standard_normal((8, 544))).astype(np.float32)  # 3 segments, large offset
    np.testing.assert_allclose(O.welch_noise_estimate_explicit(y), O.welch_noise_estimate(y), rtol=2e-5)


@pytest.mark.parametrize(
    "n,b,count,tail",
    [(512, 20, 51, [490, 492]), (512, 32, 31, [464, 480]), (1024, 40, 51, [980, 984]), (150, 32, 9, [112, 118]),
     (150, 28, 10, [112, 122]), (150, 40, 7, [100, 110]), (24, 24, 1, [0])],
)
def test_tile_starts(n, b, count, tail):
    s = O.tile_starts(n, b)
    assert len(s) == count and s[-len(tail):] == tail and s[0] == 0


def test_pyramid_weights():
    w = O.pyramid_weights(20, 20)
    assert w.dtype == np.float32 and w.max() == 10 and np.all(w[0] == 1) and np.all(w[:, 0] == 1)
    assert [w[i, i] for i in range(10)] == list(range(1, 11)) and w[10, 10] == 10
    np.testing.assert_array_equal(w, w[::-1, :])
    np.testing.assert_array_equal(w, w[:, ::-1])
    w2 = O.pyramid_weights(12, 20)
    assert w2.shape == (12, 20) and w2.max() == 6


def test_filter_by_failures():
    f = lambda a, m: O.filter_by_failures(np.array(a, bool), m).astype(int).tolist()
    assert f([1, 1, 0, 1, 1], 1) == [1, 1, 1, 0, 0]
    assert f([0, 1, 1], 1) == [1, 0, 0]
    assert f([0, 0, 0], 1) == [1, 0, 0]
    assert f([0, 0, 0], 2) == [1, 1, 0]
    assert f([0, 1, 0, 0, 1], 2) == [1, 1, 1, 1, 0]
    assert f([1, 1, 1], 1) == [1, 1, 1]


def test_roughness_closed_forms():
    ii, jj = np.mgrid[0:8, 0:6].astype(np.float32)
    ramp = 2 * ii + 3 * jj + 1
    expect = ((7 * 6) * 2 + (8 * 5) * 3) / (7 * 6 + 8 * 5) / ramp.mean()
    assert abs(O.spatial_roughness_stat(ramp) - expect) < 1e-6
    assert O.spatial_roughness_stat(np.ones((5, 5), np.float32)) == 0
    lin = np.arange(50, dtype=np.float32) + 1
    assert O.temporal_roughness_stat(lin) == 0
    alt = np.array([1, -1] * 10, np.float32)
    assert abs(O.temporal_roughness_stat(alt) - 4.0) < 1e-6
    assert np.isnan(O.temporal_roughness_stat(np.zeros(10, np.float32)))


def test_pooling_and_time_average():
    x = np.arange(4 * 6 * 20, dtype=np.float32).reshape(4, 6, 20)
    p = O.downsample_average_pooling(x, 2)
    np.testing.assert_allclose(p, x.reshape(2, 2, 3, 2, 20).mean(axis=(1, 3)), rtol=1e-6)
    xo = np.ones((5, 5, 3), np.float32)
    assert O.downsample_average_pooling(xo, 2).shape == (3, 3, 3)
    np.testing.assert_allclose(O.downsample_average_pooling(xo, 2), 1.0)
    ta = np.mean(np.reshape(p, (6, 10, 2), order="F"), axis=1)
    np.testing.assert_allclose(ta, np.reshape(p, (6, 20), order="F").reshape(6, 2, 10).mean(axis=2), rtol=1e-6)


def test_frame_batches():
    assert O.frame_batches(20000, 10000) == [(0, 20000)]
    assert O.frame_batches(1300, 256) == [(0, 256), (256, 512), (512, 768), (768, 1024), (1024, 1300)]
    assert O.frame_batches(100, 1000) == [(0, 100)]


def test_guards():
    with pytest.raises(ValueError):
        O.check_fov_size((9, 50))
    with pytest.raises(ValueError):
        O.update_block_sizes([8, 20], (100, 100))
    assert O.update_block_sizes([32, 32], (24, 20)) == [24, 20]
    with pytest.raises(ValueError):
        O.window_chunk_candidates(5000, 1000, 5000)
    with pytest.raises(ValueError):
        O.window_chunk_candidates(100, 1000, 200)
    av, n = O.window_chunk_candidates(600, 1300, 600)
    assert av.tolist() == [0, 600, 700] and n == 1


def test_exact_lowrank_block_rank_rule():
    """Smooth exact rank-3 block + small noise: recovered rank = 3 (+1 kept first failure)."""
    rng = np.random.default_rng(3)
    ii, jj = np.mgrid[0:20, 0:20]
    sp_ = np.stack([np.exp(-((ii - a) ** 2 + (jj - b) ** 2) / 30.0) for a, b in [(5, 5), (14, 8), (9, 15)]], 2)
    tt = np.cumsum(rng.standard_normal((3, 1000)), axis=1)
    block = (np.tensordot(sp_, tt, (2, 0)) * 5 + rng.standard_normal((20, 20, 1000))).astype(np.float32)
    sk = rng.standard_normal((100, 20)).astype(np.float32)
    u, good, v, _ = O.single_block_md(block, sk, 10, 10, 2, 1.35, 2.3)
    keep = O.filter_by_failures(good > 0, 1)
    assert keep.sum() == 4 and good[:3].tolist() == [1, 1, 1]
    np.testing.assert_allclose(np.reshape(u, (400, 10), order="F").T @ np.reshape(u, (400, 10), order="F"), np.eye(10), atol=1e-4)
