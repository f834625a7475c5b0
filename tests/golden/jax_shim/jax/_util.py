import numpy as np


def to_x32(x):
    """x64-disabled dtype canonicalisation."""
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            return x.astype(np.float32)
        if x.dtype == np.int64:
            return x.astype(np.int32)
        return x
    if isinstance(x, np.float64):
        return np.float32(x)
    if isinstance(x, tuple):
        return tuple(to_x32(y) for y in x)
    return x
