import numpy as _np

ArrayLike = _np.ndarray
