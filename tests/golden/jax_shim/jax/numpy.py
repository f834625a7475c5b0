"""jax.numpy stand-in: every call is numpy's, with float64/int64 inputs and outputs canonicalised
to 32 bit (JAX with x64 disabled)."""
import sys
import types

import numpy as _np

from ._util import to_x32 as _to_x32

ndarray = _np.ndarray
float32 = _np.float32
int32 = _np.int32
pi = _np.pi
newaxis = None


def _wrap(f):
    def g(*args, **kwargs):
        args = [_to_x32(a) if isinstance(a, (_np.ndarray, _np.float64)) else a for a in args]
        kwargs = {k: (_to_x32(v) if isinstance(v, (_np.ndarray, _np.float64)) else v) for k, v in kwargs.items()}
        out = f(*args, **kwargs)
        if isinstance(out, tuple) and not hasattr(out, "_fields"):
            return tuple(_to_x32(o) for o in out)
        if hasattr(out, "_fields"):  # namedtuple results (qr / svd in numpy 2)
            return tuple(_to_x32(_np.asarray(o)) for o in out)
        return _to_x32(out)

    g.__name__ = getattr(f, "__name__", "wrapped")
    return g


class _Linalg(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _wrap(getattr(_np.linalg, name))


linalg = _Linalg("jax.numpy.linalg")
sys.modules["jax.numpy.linalg"] = linalg


def array(x, dtype=None):
    return _to_x32(_np.array(x, dtype=dtype))


def asarray(x, dtype=None):
    return _to_x32(_np.asarray(x, dtype=dtype))


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _wrap(getattr(_np, name))
