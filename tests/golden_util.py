"""Helpers that turn a golden fixture (tests/golden/case_*.npz) back into inputs: the movie, the
reference's parameters and a `Draws` object holding the identical random numbers."""
import json
import os

import numpy as np

from oracle.pmd_oracle import Draws, normal_from_seed, tile_starts, update_block_sizes
from synth import make_movie

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_NAMES = ["main_F", "prune_C_u16", "tiny_noNorm", "wide_R", "windows"]


def load_case(name):
    g = np.load(os.path.join(GOLDEN_DIR, "case_%s.npz" % name), allow_pickle=False)
    spec = json.loads(str(g["spec"]))
    mv = dict(spec["movie"])
    dtype = np.dtype(mv.pop("dtype"))
    movie = make_movie(dtype=dtype, **mv)
    chk = g["movie_checksum"]
    assert abs(float(movie.astype(np.float64).sum()) - chk[0]) <= 1e-6 * abs(chk[0]), "synthetic movie drifted"
    assert float(movie[3, 5, 7]) == chk[1]
    return g, spec, movie


def draws_from_case(g, spec, movie, lazy_sim=False):
    """Rebuild every random draw in the reference's consumption order (see make_golden.py)."""
    kw = spec["kwargs"]
    T, d1, d2 = movie.shape
    seeds, shapes = g["normal_seeds"], g["normal_shapes"]

    def normal(i):
        shp = tuple(int(x) for x in shapes[i] if x >= 0)
        return normal_from_seed(int(seeds[i]), shp)

    pos = 0
    ci = 0
    d = Draws()
    bg_rank = kw.get("background_rank", 15)
    if bg_rank > 0:
        d.bg_frames = g["choice_%d" % ci].tolist()
        ci += 1
        d.bg_sketch = normal(pos)
        pos += 1
    frame_range = spec["frame_range"]
    wc = kw.get("window_chunks", None) or frame_range
    if T >= frame_range:
        wc = min(wc, frame_range)
        starts = np.sort(g["choice_%d" % ci])
        ci += 1
        fr = []
        for k in starts:
            fr.extend(range(int(k), int(min(k + wc, T))))
        d.init_frames = fr
    # 250 simulations: (noise, sketch) pairs
    n_after_thr = int(g["n_log_after_thr"])
    n_sim = (n_after_thr - pos) // 2
    sim_idx = [(pos + 2 * i, pos + 2 * i + 1) for i in range(n_sim)]
    if lazy_sim:
        d.sim_noise = _LazyList([a for a, _ in sim_idx], normal)
        d.sim_sketch = _LazyList([b for _, b in sim_idx], normal)
    else:
        d.sim_noise = [normal(a) for a, _ in sim_idx]
        d.sim_sketch = [normal(b) for _, b in sim_idx]
    pos = n_after_thr
    bs = []
    for n in g["block_n_sketch"]:
        bs.append([normal(pos + i) for i in range(int(n))])
        pos += int(n)
    d.block_sketches = bs
    if kw.get("rank_prune", False):
        d.prune_sketch = normal(pos)
        pos += 1
    assert pos == len(seeds), (pos, len(seeds))
    return d


class _LazyList:
    def __init__(self, idx, fn):
        self.idx, self.fn = idx, fn

    def __len__(self):
        return len(self.idx)

    def __iter__(self):
        for i in self.idx:
            yield self.fn(i)

    def __getitem__(self, i):
        return self.fn(self.idx[i])


def block_grid(spec, movie):
    T, d1, d2 = movie.shape
    bh, bw = update_block_sizes(spec["block_sizes"], (d1, d2))
    return [(k, j) for k in tile_starts(d1, bh) for j in tile_starts(d2, bw)], (bh, bw)
