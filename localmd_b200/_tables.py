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


# ------------------------------------------------------------------------------------------------------------------------
# K1 on the tensor cores (csrc/stats_tc.cu)
# ------------------------------------------------------------------------------------------------------------------------
def _tf32_trunc(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _bf16_bits(x):
    """round-to-nearest-even bfloat16 bit patterns (uint32, low 16 bits) of float32 values."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint32)


def welch_dif_matrices():
    """The two real (64, 128) matrices of the radix-2 decimation-in-frequency form of the segment transform, float64.
    With a[t] = w[t] x[t] + (1 - w[t]) x[t + 128] and b[t] = w[t] x[t] - (1 - w[t]) x[t + 128]  (w = periodic Hann of 256
    points, w[t + 128] = 1 - w[t]; t = 0..127), the windowed 256-point DFT is
        X[f] = sum_t a[t] e^{-2 pi i f t / 256}   for even f,      X[f] = sum_t b[t] e^{-2 pi i f t / 256}   for odd f.
    Rows 2 g, 2 g + 1 = (cos, -sin) of bin f = 66 + 2 g (even matrix; bin 128 scaled by sqrt(1/2): the one-sided density
    does not double the Nyquist bin) and of bin f = 65 + 2 g (odd matrix)."""
    t = np.arange(128, dtype=np.float64)

    def rows(bins):
        out = []
        for f in bins:
            sc = np.sqrt(0.5) if f == 128 else 1.0
            out.append(sc * np.cos(2.0 * np.pi * f * t / 256.0))
            out.append(-sc * np.sin(2.0 * np.pi * f * t / 256.0))
        return np.array(out)

    return rows(range(66, 129, 2)), rows(range(65, 128, 2))


def welch_tc_tables():
    """Operand image of the tensor-core stats kernel as raw bytes (uint8, 131072 + 512):
      [even | odd] TF32 parts, then [even | odd] bf16 pair parts, each 4 K atoms of 32 frames x (64 rows x 128 bytes) in
      the canonical K-major SWIZZLE_128B shared-memory layout (16-byte piece c of row n at (n >> 3) * 1024 + (n & 7) * 128
      + ((c ^ (n & 7)) << 4));  TF32 part = value with the low 13 mantissa bits cleared, pair part per element
      bf16(lo) | bf16(hi) << 16 (lo = value - TF32 part);  then w[0..127] (periodic Hann, first half) as float32."""
    be, bo = welch_dif_matrices()
    img = np.zeros(131072 // 4, dtype=np.uint32)
    for par, mat in enumerate((be, bo)):
        m32 = mat.astype(np.float32)
        hi = _tf32_trunc(m32)
        lo = (m32 - hi).astype(np.float32)
        pair = _bf16_bits(lo) | (_bf16_bits(hi) << np.uint32(16))
        n = np.arange(64)[:, None]
        k = np.arange(128)[None, :]
        ka, kk = k // 32, k % 32
        c, j = kk // 4, kk % 4
        off_bytes = ka * 8192 + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4) + 4 * j
        idx = (off_bytes // 4).astype(np.int64)
        img[par * 8192 + idx] = hi.view(np.uint32)
        img[16384 + par * 8192 + idx] = pair
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(128, dtype=np.float64) / 256.0)).astype(np.float32)
    return np.concatenate([img.view(np.uint8), w.view(np.uint8)])


def welch_dif_reference(x):
    """NumPy evaluation (float64) of what the tensor-core kernel computes for one chunk; the CPU tests prove the
    formulation against scipy.signal.welch."""
    x = np.asarray(x, dtype=np.float64)
    npts = x.shape[-1]
    nseg = (npts - 128) // 128 if npts >= 256 else 0
    if nseg == 0:
        return np.zeros(x.shape[:-1])
    be, bo = welch_dif_matrices()
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(128) / 256.0)
    x = x - x[..., :1]
    tot = np.zeros(x.shape[:-1])
    for s in range(nseg):
        p = w * x[..., 128 * s : 128 * s + 128]
        q = (1.0 - w) * x[..., 128 * s + 128 : 128 * s + 256]
        tot += (((p + q) @ be.T) ** 2).sum(axis=-1) + (((p - q) @ bo.T) ** 2).sum(axis=-1)
    return np.sqrt(tot / (64.0 * 96.0 * nseg))
