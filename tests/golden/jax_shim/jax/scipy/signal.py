"""jax.scipy.signal.welch -> scipy.signal.welch (identical defaults: fs=1, hann, nperseg=256,
constant detrend, one-sided density, mean over segments)."""
import numpy as _np
import scipy.signal as _ss


def welch(x, **kw):
    f, p = _ss.welch(_np.asarray(x, dtype=_np.float32), **kw)
    return f.astype(_np.float32), p.astype(_np.float32)
