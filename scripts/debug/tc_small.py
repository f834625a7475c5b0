"""Small single-shot run of the tensor-core projection kernel (debug helper)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from localmd_b200 import ops
from localmd_b200.decomposition import tile_starts
d1, d2, bh, bw, K, T = 60, 96, 20, 20, 3, 300
rng = np.random.default_rng(0)
rows, cols = tile_starts(d1, bh), tile_starts(d2, bw)
nb = len(rows) * len(cols)
ranks = rng.integers(1, 5, nb).astype(np.int32)
col0 = np.concatenate([[0], np.cumsum(ranks)[:-1]]).astype(np.int64)
n_local = int(ranks.sum())
dev = torch.device("cuda")
uv = torch.randn((n_local, bh * bw), device=dev)
bg = torch.randn((K, d1 * d2), device=dev)
movie = torch.randn((T, d1 * d2), device=dev)
st = ops.make_strips_tc(rows, cols, bh, bw, d1, d2, ranks, col0, K)
print("items", st["items"])
std = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in st.items()}
bimg = ops.pack_strips_tc(std, uv, bg, bh * bw, d2)
z = torch.zeros((n_local + K, T), device=dev)
ops.project_stream_tc(movie, d2, std, bimg, None, None, z[:n_local], z[n_local:])
torch.cuda.synchronize()
ref = bg.double() @ movie.double().t()
print("bg err", ((z[n_local:].double() - ref).abs().max() / ref.abs().max()).item())
