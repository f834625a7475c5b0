"""Deterministic synthetic functional-imaging movies for the tests (NumPy, platform-stable PCG64).

Y = mu + sigma_px * ( A.C + Bg.F + eps ):  A Gaussian blobs, C spike trains convolved with an
exponential, Bg.F a smooth low-rank background, eps ~ N(0,1), sigma_px ~ U(0.5,2), mu ~ U(100,300)
(the small-scale twin of the generator in SURVEY.md section 8d / localmd_b200/synthetic.py)."""
import numpy as np


def make_movie(T, d1, d2, n_cells=6, seed=0, dtype=np.float32, bg_rank=2, noise=1.0, blob_sigma=(1.5, 3.0)):
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:d1, 0:d2].astype(np.float64)
    a = np.zeros((n_cells, d1, d2))
    for c in range(n_cells):
        cy, cx = rng.uniform(0, d1), rng.uniform(0, d2)
        s = rng.uniform(*blob_sigma)
        a[c] = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s)) * rng.uniform(2, 6)
    spikes = (rng.uniform(size=(n_cells, T)) < 0.02).astype(np.float64)
    kern = np.exp(-np.arange(60) / 15.0)
    c_tr = np.stack([np.convolve(spikes[c], kern)[:T] for c in range(n_cells)]) if n_cells else np.zeros((0, T))
    sig = np.tensordot(c_tr.T, a, axes=(1, 0)) if n_cells else np.zeros((T, d1, d2))
    for b in range(bg_rank):
        img = np.cos(np.pi * (b + 1) * yy / d1) * np.cos(np.pi * (b + 0.5) * xx / d2)
        walk = np.cumsum(rng.standard_normal(T)) * 0.05
        sig = sig + walk[:, None, None] * img[None]
    eps = rng.standard_normal((T, d1, d2)) * noise
    sigma_px = rng.uniform(0.5, 2.0, size=(d1, d2))
    mu = rng.uniform(100, 300, size=(d1, d2))
    y = mu[None] + sigma_px[None] * (sig + eps)
    if np.issubdtype(np.dtype(dtype), np.integer):
        y = np.clip(np.rint(y), np.iinfo(dtype).min, np.iinfo(dtype).max)
    return y.astype(dtype)
