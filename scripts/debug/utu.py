import sys, numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import oracle.pmd_oracle as O
from localmd_b200 import ops
from localmd_b200.decomposition import SparseU
from test_kernels_gpu import _random_sparse_u, dev
rng = np.random.default_rng(132)
bh, bw, d1, d2, max_rank, K = 20, 20, 112, 95, 5, 3
starts, ranks, col0, uv, bg, U = _random_sparse_u(rng, d1, d2, bh, bw, max_rank, K)
su = SparseU(starts, dev(starts), bh, bw, d1, d2, ranks.astype(np.int64), dev(ranks), dev(uv.astype(np.float64)), dev(uv), dev(bg))
(rowptr, cols, vals), c = su.gram()
nl = int(ranks.sum())
G = (U.T @ U).toarray()
L = sp.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), rowptr.cpu().numpy()), shape=(nl, nl)).toarray()
print("L err", np.abs(L - G[:nl, :nl]).max())
print("C err", np.abs(c.cpu().numpy() - G[:, nl:]).max())
right = rng.standard_normal((U.shape[1], 9))
r = dev(right)
rows = torch.arange(nl, dtype=torch.int32, device='cuda')
lr = ops.reconstruct_f64(rowptr, cols, vals, r[:nl].contiguous(), rows).t().cpu().numpy()
print("L r err", np.abs(lr - G[:nl, :nl] @ right[:nl]).max())
got = su.utu_times_f64(r).cpu().numpy()
print("total err", np.abs(got - G @ right).max())
