"""jax.random stand-in.  normal(key, shape) is a documented function of the integer seed:
numpy.random.Generator(PCG64(seed mod 2**32)).standard_normal(shape, dtype=float32).
Every call is appended to RANDOM_LOG as (seed, shape)."""
import numpy as _np

RANDOM_LOG = []


def PRNGKey(seed):
    return _np.array([int(seed) & 0xFFFFFFFF], dtype=_np.uint32)


def normal_from_seed(seed, shape):
    rng = _np.random.Generator(_np.random.PCG64(int(seed) & 0xFFFFFFFF))
    return rng.standard_normal(tuple(int(s) for s in shape), dtype=_np.float32)


def normal(key, shape, dtype=_np.float32):
    seed = int(_np.asarray(key).ravel()[0])
    RANDOM_LOG.append((seed, tuple(int(s) for s in shape)))
    return normal_from_seed(seed, shape)
