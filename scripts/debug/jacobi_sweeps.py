"""Histogram of the Jacobi sweeps the block stage really uses (development build: PMD_LIB_PATH=localmd_b200/libpmd_tune.so,
compiled with -DPMD_TUNE).  Runs one C2-like decomposition (fewer frames) and prints the histogram per precision."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import localmd_b200  # noqa: E402
from localmd_b200 import _lib  # noqa: E402
from localmd_b200.synthetic import make_movie  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda")
movie = make_movie(T, 512, 512, n_cells=400, seed=1234, device=dev)
L = _lib.lib()
out = (ctypes.c_ulonglong * 128)()
for it in range(2):
    L.pmd_debug_jacobi_hist(out, 1)
    localmd_b200.localmd_decomposition(movie, block_sizes=[20, 20], frame_range=5000, rank_prune=True, seed=0)
L.pmd_debug_jacobi_hist(out, 1)
h = np.array(list(out), dtype=np.int64).reshape(2, 64)
for name, row in (("float64", h[0]), ("float32", h[1])):
    nz = np.nonzero(row)[0]
    print(name, {int(i): int(row[i]) for i in nz}, "mean %.2f" % ((row * np.arange(64)).sum() / max(1, row.sum())))
