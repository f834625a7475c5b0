"""NumPy-backed stand-in for the handful of JAX entry points apasarkar/localmd uses.

TEST INFRASTRUCTURE ONLY.  jax/jaxlib are not installable in this image (no network), so the
golden-fixture generator (tests/golden/make_golden.py) imports the UNMODIFIED reference package
from /root/reference with this directory first on sys.path.  Every primitive is mapped to its
NumPy/SciPy float32 equivalent ("x64 disabled" semantics: float64 results are cast to float32 at
every jnp call and at every jit boundary).  The reference's own Python control flow, reshape
orders, index conventions and host-side scipy.sparse code therefore execute as written; only the
XLA numerics are substituted (LAPACK via NumPy instead of LAPACK via XLA-CPU).

Every Gaussian draw goes through jax.random.normal below, which is a *documented* function of
the integer seed (numpy Generator PCG64 seeded with seed mod 2**32, standard_normal float32) and
is logged in RANDOM_LOG so the fixtures can store (seed, shape) pairs instead of the matrices.
"""
import functools

import numpy as _np

from . import numpy as numpy  # noqa: F401  (jax.numpy)
from . import lax as lax  # noqa: F401
from . import random as random  # noqa: F401
from . import scipy as scipy  # noqa: F401
from . import typing as typing  # noqa: F401
from ._util import to_x32 as _to_x32

Array = _np.ndarray


def _convert_arg(a):
    # jit boundary: device_put with x64 disabled -> float32 / int32 copies (JAX arrays are immutable,
    # so in-place ops inside a jitted function must never alias the caller's buffers)
    if isinstance(a, _np.ndarray):
        return _to_x32(_np.array(a, copy=True))
    if isinstance(a, (_np.floating,)):
        return _np.float32(a)
    if isinstance(a, (list, tuple)) and len(a) and all(isinstance(x, _np.ndarray) for x in a):
        return type(a)(_convert_arg(x) for x in a)
    return a


def jit(fun=None, static_argnums=None, **_kw):
    if fun is None:
        return functools.partial(jit, static_argnums=static_argnums)
    if static_argnums is None:
        static = ()
    elif isinstance(static_argnums, int):
        static = (static_argnums,)
    else:
        static = tuple(static_argnums)

    @functools.wraps(fun)
    def wrapped(*args, **kwargs):
        conv = [a if i in static else _convert_arg(a) for i, a in enumerate(args)]
        kconv = {k: _convert_arg(v) for k, v in kwargs.items()}
        return fun(*conv, **kconv)

    return wrapped


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) == 1 and len(args) > 1:
            axes = tuple(axes) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.asarray(a).shape[ax]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _np.take(_np.asarray(a), i, axis=ax) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if isinstance(outs[0], tuple):
            return tuple(_to_x32(_np.stack([o[j] for o in outs], axis=out_axes)) for j in range(len(outs[0])))
        return _to_x32(_np.stack(outs, axis=out_axes))

    return mapped
