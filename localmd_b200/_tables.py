"""Constant tables for the stats kernel (csrc/stats.cu).

Folded Hann-windowed DFT of a 256-sample segment, bins k = 65..128 (preprocessing_utils.py:28-37 takes
exactly these bins of welch(trace, noverlap=128)):  with w[n] = 0.5 - 0.5 cos(2 pi n / 256) = w[256-n],
    Re X[k] = sum_{n=1..128} tab_cos[n-1][k-65] * e[n],   e[n] = x[n] + x[256-n]  (e[128] = x[128])
    Im X[k] = -sum_{n=1..127} tab_sin[n-1][k-65] * o[n],  o[n] = x[n] - x[256-n]
(w[0] = 0, so the n = 0 sample never contributes)."""
import numpy as np


def welch_tables():
    n = np.arange(1, 129, dtype=np.float64)[:, None]  # 1..128
    k = np.arange(65, 129, dtype=np.float64)[None, :]  # 65..128
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / 256.0)
    ang = 2.0 * np.pi * n * k / 256.0
    tab_cos = w * np.cos(ang)
    tab_sin = w * np.sin(ang)
    tab_sin[127, :] = 0.0  # n = 128: sin(pi k) = 0 exactly
    tab_cos[127, :] = np.where((np.arange(65, 129) % 2) == 0, 1.0, -1.0)
    return np.ascontiguousarray(tab_cos, dtype=np.float32), np.ascontiguousarray(tab_sin, dtype=np.float32)


def welch_from_tables_reference(x):
    """NumPy evaluation of exactly what the kernel computes for one chunk (float64); used by the CPU
    tests to prove the folded-table identity against scipy.signal.welch."""
    x = np.asarray(x, dtype=np.float64)
    tc, ts = welch_tables()
    tc, ts = tc.astype(np.float64), ts.astype(np.float64)
    npts = x.shape[-1]
    nseg = (npts - 128) // 128 if npts >= 256 else 0
    if nseg == 0:
        return np.zeros(x.shape[:-1])
    x = x - x[..., :1]
    tot = np.zeros(x.shape[:-1])
    for s in range(nseg):
        seg = x[..., 128 * s : 128 * s + 256]
        e = seg[..., 1:129].copy()
        e[..., :127] += seg[..., 255:128:-1]
        o = seg[..., 1:129] - np.concatenate([seg[..., 255:128:-1], seg[..., 128:129]], axis=-1)
        re = e @ tc
        im = o @ ts
        pw = re**2 + im**2
        pw[..., 63] *= 0.5
        tot += pw.sum(axis=-1)
    return np.sqrt(tot / (64.0 * 96.0 * nseg))


def welch_fft_tables():
    """Tables of the FFT-based stats kernel (csrc/stats_fft.cu), 772 float32:
    hann[256] (periodic Hann), tw128[128][2] = (cos, -sin)(2 pi j / 128), tw256[130][2] = (cos, sin)(2 pi k / 256)."""
    j = np.arange(256, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * j / 256.0)
    a = 2.0 * np.pi * np.arange(128, dtype=np.float64) / 128.0
    tw128 = np.stack([np.cos(a), -np.sin(a)], axis=1)
    b = 2.0 * np.pi * np.arange(130, dtype=np.float64) / 256.0
    tw256 = np.stack([np.cos(b), np.sin(b)], axis=1)
    return np.concatenate([hann, tw128.reshape(-1), tw256.reshape(-1)]).astype(np.float32)


def welch_fft_reference(x):
    """NumPy evaluation (float64) of exactly what the FFT kernel computes for one chunk: per 256-sample segment a
    complex 128-point FFT of z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1], the real-FFT split, and the power of bins
    65..128.  Used by the CPU tests to prove the formulation against scipy.signal.welch."""
    x = np.asarray(x, dtype=np.float64)
    npts = x.shape[-1]
    nseg = (npts - 128) // 128 if npts >= 256 else 0
    if nseg == 0:
        return np.zeros(x.shape[:-1])
    tab = welch_fft_tables().astype(np.float64)
    hann, tw256 = tab[:256], tab[512:].reshape(130, 2)
    x = x - x[..., :1]
    tot = np.zeros(x.shape[:-1])
    for s in range(nseg):
        seg = x[..., 128 * s : 128 * s + 256] * hann
        z = np.fft.fft(seg[..., 0::2] + 1j * seg[..., 1::2], axis=-1)  # Z[0..127]
        for k in range(65, 128):
            A, B = z[..., k], np.conj(z[..., 128 - k])
            wk = tw256[k, 0] - 1j * tw256[k, 1]
            X = 0.5 * (A + B) - 0.5j * wk * (A - B)
            tot += np.abs(X) ** 2
        tot += 0.5 * (z[..., 0].real - z[..., 0].imag) ** 2
    return np.sqrt(tot / (64.0 * 96.0 * nseg))
