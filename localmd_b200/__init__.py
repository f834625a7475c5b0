"""localmd_b200: B200-native (sm_100a) implementation of the hot path of apasarkar/localmd behind the
reference's own Python API (localmd/__init__.py:1-7 exports the same names)."""
from .dataset import TiffArray, lazy_data_loader  # noqa: F401
from .decomposition import compute_lowrank_factorized_svd, localmd_decomposition, projected_svd  # noqa: F401
from .io import load_npz, save_npz  # noqa: F401
from .pmdarray import PMDArray  # noqa: F401

__version__ = "0.1.0"
