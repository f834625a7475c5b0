"""jax.experimental.sparse.BCOO stand-in: a float32 scipy CSR that supports `@ dense`."""
import numpy as _np
import scipy.sparse as _sp


class BCOO:
    def __init__(self, mat):
        self.mat = _sp.csr_matrix(mat).astype(_np.float32)
        self.shape = self.mat.shape

    @classmethod
    def from_scipy_sparse(cls, mat):
        return cls(mat)

    def __matmul__(self, other):
        return _np.asarray(self.mat @ _np.asarray(other, dtype=_np.float32), dtype=_np.float32)
