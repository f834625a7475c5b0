"""jax.lax stand-in (dynamic_slice, reduce_window(add), select, add)."""
import numpy as _np

from ._util import to_x32 as _to_x32


def add(a, b):
    return a + b


def dynamic_slice(operand, start_indices, slice_sizes):
    operand = _np.asarray(operand)
    idx = []
    for dim, (s, n) in enumerate(zip(start_indices, slice_sizes)):
        if n > operand.shape[dim]:
            raise TypeError("dynamic_slice: slice size %d larger than operand dim %d" % (n, operand.shape[dim]))
        s = int(min(max(int(s), 0), operand.shape[dim] - n))  # XLA clamps the start index
        idx.append(slice(s, s + n))
    return operand[tuple(idx)]


def select(pred, on_true, on_false):
    return _np.where(pred, on_true, on_false)


def reduce_window(operand, init_value, computation, window_dimensions, window_strides, padding):
    """Sum-pooling reduce_window with XLA 'SAME' / 'VALID' padding (only `add` is needed)."""
    assert computation is add
    x = _to_x32(_np.asarray(operand))
    nd = x.ndim
    if isinstance(padding, str):
        pads = []
        for d in range(nd):
            n, w, s = x.shape[d], window_dimensions[d], window_strides[d]
            if padding.upper() == "SAME":
                out = -(-n // s)
                total = max((out - 1) * s + w - n, 0)
                pads.append((total // 2, total - total // 2))
            else:
                pads.append((0, 0))
    else:
        pads = list(padding)
    xp = _np.pad(x, pads, mode="constant", constant_values=init_value)
    out_shape = [(xp.shape[d] - window_dimensions[d]) // window_strides[d] + 1 for d in range(nd)]
    out = _np.full(out_shape, init_value, dtype=x.dtype)
    import itertools

    for offs in itertools.product(*[range(w) for w in window_dimensions]):
        sl = tuple(
            slice(o, o + (out_shape[d] - 1) * window_strides[d] + 1, window_strides[d]) for d, o in enumerate(offs)
        )
        out = out + xp[sl]
    return out
