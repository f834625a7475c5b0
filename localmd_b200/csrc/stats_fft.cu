// K1 (v2): per-pixel mean + Welch high-band noise estimate in one streaming pass over the movie, with the
// segment spectra computed by an in-register / shared-memory FFT.
//
// Replaces pmd_loader.py:203-291 and preprocessing_utils.py:10-40 of the reference: per 1024-frame chunk and
// pixel, scipy/jax `welch(trace, noverlap=128)` (Hann-periodic window, 256-sample segments, hop 128, constant
// detrend, one-sided density, mean over segments), then sqrt(mean(0.5 * Pxx[65..128])).
//   * The Hann-windowed DFT of a constant is non-zero only at bins 0 and +-1, so the per-segment mean removal
//     cannot change bins >= 65; a per-pixel offset (the chunk's first frame) is subtracted only for rounding.
//   * A real 256-point DFT is one complex 128-point FFT of z[n] = x[2n] + i x[2n+1] plus the split
//         X[k] = (Z[k] + conj Z[128-k]) / 2 - i W256^k (Z[k] - conj Z[128-k]) / 2,
//     and the 128-point FFT is done as 16 x 8 (Cooley-Tukey, n = 8 n1 + n2, k = k1 + 16 k2):
//         pass A: warp n2 does the 16-point FFT over n1 in registers and applies the twiddle W128^(n2 k1);
//         pass B: warp k1 (and k1 + 8) does the 8-point FFT over n2, in place;
//         pass C: warp w evaluates |X[k]|^2 for k = 65 + w + 8 j.
//     ~4.5 kflop per segment and pixel instead of 64 x 128 x 2 multiply-adds for the direct transform.
// One CTA = 32 pixels (one per lane: every shared-memory access is conflict free) x one 1024-frame chunk;
// 8 warps; a 256-frame ring buffer; the next 128 frames are in flight while a segment is transformed.
//   sigma^2 = 1/(64*96*nseg) * sum_seg ( sum_{k=65..127} |X[k]|^2 + 0.5 |X[128]|^2 )      (96 = sum w^2)
#include "common.cuh"

namespace pmd {

constexpr int kSFPix = 32;
constexpr int kSFThreads = 256;
constexpr int kSFChunk = 1024;
constexpr int kSFHop = 128;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// forward 4-point DFT in place: (x0, x1, x2, x3) -> (X0, X1, X2, X3)
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s0 = cadd(x0, x2), s1 = csub(x0, x2), s2 = cadd(x1, x3), s3 = mul_mi(csub(x1, x3));
    x0 = cadd(s0, s2);
    x2 = csub(s0, s2);
    x1 = cadd(s1, s3);
    x3 = csub(s1, s3);
}

// forward 16-point FFT: a[n] -> a[k] (natural order in, natural order out)
__device__ __forceinline__ void fft16(float2 (&a)[16]) {
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    // step 1: for each r, 4-point DFT over m of a[4m + r]  -> T_r[q] stored at a[4q + r]
#pragma unroll
    for (int r = 0; r < 4; ++r) dft4(a[r], a[4 + r], a[8 + r], a[12 + r]);
    // step 2: T_r[q] *= W16^(r q)
    a[4 + 1] = cmul(a[4 + 1], make_float2(c1, -s1));    // q=1 r=1 : W^1
    a[4 + 2] = cmul(a[4 + 2], make_float2(h, -h));      // q=1 r=2 : W^2
    a[4 + 3] = cmul(a[4 + 3], make_float2(s1, -c1));    // q=1 r=3 : W^3
    a[8 + 1] = cmul(a[8 + 1], make_float2(h, -h));      // q=2 r=1 : W^2
    a[8 + 2] = mul_mi(a[8 + 2]);                        // q=2 r=2 : W^4 = -i
    a[8 + 3] = cmul(a[8 + 3], make_float2(-h, -h));     // q=2 r=3 : W^6
    a[12 + 1] = cmul(a[12 + 1], make_float2(s1, -c1));  // q=3 r=1 : W^3
    a[12 + 2] = cmul(a[12 + 2], make_float2(-h, -h));   // q=3 r=2 : W^6
    a[12 + 3] = cmul(a[12 + 3], make_float2(-c1, s1));  // q=3 r=3 : W^9
    // step 3: for each q, 4-point DFT over r -> Y[q + 4 s] ; currently T_r[q] sits at a[4q + r]
    float2 y[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float2 t0 = a[4 * q], t1 = a[4 * q + 1], t2 = a[4 * q + 2], t3 = a[4 * q + 3];
        dft4(t0, t1, t2, t3);
        y[q] = t0;
        y[q + 4] = t1;
        y[q + 8] = t2;
        y[q + 12] = t3;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = y[i];
}

// forward 8-point FFT
__device__ __forceinline__ void fft8(float2 (&y)[8]) {
    constexpr float h = 0.70710678118654752f;
    float2 e0 = y[0], e1 = y[2], e2 = y[4], e3 = y[6];
    float2 o0 = y[1], o1 = y[3], o2 = y[5], o3 = y[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = cmul(o1, make_float2(h, -h));
    o2 = mul_mi(o2);
    o3 = cmul(o3, make_float2(-h, -h));
    y[0] = cadd(e0, o0); y[4] = csub(e0, o0);
    y[1] = cadd(e1, o1); y[5] = csub(e1, o1);
    y[2] = cadd(e2, o2); y[6] = csub(e2, o2);
    y[3] = cadd(e3, o3); y[7] = csub(e3, o3);
}

template <typename T>
__global__ void __launch_bounds__(kSFThreads, 3)
stats_fft_kernel(const T* __restrict__ movie, int64_t t_local, int64_t d, double inv_total, const float* __restrict__ tab,
                 float* __restrict__ mean_part, float* __restrict__ noise_part) {
    extern __shared__ __align__(16) float sfm[];
    float* ring = sfm;                                            // [256][32]
    float2* zb = reinterpret_cast<float2*>(ring + 256 * kSFPix);  // [128][32]
    float* hann = reinterpret_cast<float*>(zb + 128 * kSFPix);    // [256]
    float2* tw128 = reinterpret_cast<float2*>(hann + 256);        // [128]  (cos, -sin)(2 pi j / 128)
    float2* tw256 = tw128 + 128;                                  // [130]  (cos, sin)(2 pi k / 256)

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t px = (int64_t)blockIdx.x * kSFPix + lane;
    const bool valid = px < d;
    const int chunk = blockIdx.y;
    const int64_t f_begin = (int64_t)chunk * kSFChunk;
    const int n = (int)min((int64_t)kSFChunk, t_local - f_begin);
    const int nseg = n >= 256 ? (n - kSFHop) / kSFHop : 0;
    const int nhb = (n + kSFHop - 1) / kSFHop;

    for (int i = tid; i < 256 + 256 + 260; i += kSFThreads) hann[i] = tab[i];

    const T* col = movie + f_begin * d + (valid ? px : 0);
    const float c0 = valid ? to_f32(col[0]) : 0.f;
    double msum = 0.0;
    float pw = 0.f;

    T pre[16];
    auto prefetch = [&](int hb) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int fr = hb * kSFHop + w + 8 * i;
            pre[i] = (valid && fr < n) ? col[(int64_t)fr * d] : T(0);
        }
    };
    auto commit = [&](int hb) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int fr = hb * kSFHop + w + 8 * i;
            float v = 0.f;
            if (valid && fr < n) {
                const float x = to_f32(pre[i]);
                msum += (double)x;
                v = x - c0;
            }
            ring[(fr & 255) * kSFPix + lane] = v;
        }
    };

    prefetch(0);
    commit(0);
    for (int hb = 0; hb < nhb; ++hb) {
        if (hb + 1 < nhb) prefetch(hb + 1);
        __syncthreads();  // half-blocks hb-1 and hb are in the ring (and the tables, first time round)
        const int s = hb - 1;
        if (s >= 0 && s < nseg) {
            const int base = s * kSFHop;
            // ---- pass A: n2 = w ; 16-point FFT over n1, twiddle, store Y[n2][k1]
            {
                float2 a[16];
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) {
                    const int j = 16 * n1 + 2 * w;
                    a[n1].x = ring[((base + j) & 255) * kSFPix + lane] * hann[j];
                    a[n1].y = ring[((base + j + 1) & 255) * kSFPix + lane] * hann[j + 1];
                }
                fft16(a);
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    const float2 y = k1 == 0 ? a[0] : cmul(a[k1], tw128[(w * k1) & 127]);
                    zb[(w * 16 + k1) * kSFPix + lane] = y;
                }
            }
            __syncthreads();
            // ---- pass B: k1 = w, w + 8 ; 8-point FFT over n2, in place -> Z[k1 + 16 k2]
#pragma unroll
            for (int rep = 0; rep < 2; ++rep) {
                const int k1 = w + 8 * rep;
                float2 y[8];
#pragma unroll
                for (int n2 = 0; n2 < 8; ++n2) y[n2] = zb[(n2 * 16 + k1) * kSFPix + lane];
                fft8(y);
#pragma unroll
                for (int k2 = 0; k2 < 8; ++k2) zb[(k2 * 16 + k1) * kSFPix + lane] = y[k2];
            }
            __syncthreads();
            // ---- pass C: |X[k]|^2, k = 65 + w + 8 j
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = 65 + w + 8 * j;
                if (k == 128) {
                    const float2 z0 = zb[lane];
                    const float x = z0.x - z0.y;
                    pw = fmaf(0.5f * x, x, pw);
                } else {
                    const float2 A = zb[k * kSFPix + lane], B = zb[(128 - k) * kSFPix + lane];
                    const float sr = A.x + B.x, si = A.y - B.y, dr = A.x - B.x, di = A.y + B.y;
                    const float2 cs = tw256[k];
                    const float xr = 0.5f * (sr - cs.y * dr + cs.x * di);
                    const float xi = 0.5f * (si - cs.y * di - cs.x * dr);
                    pw = fmaf(xr, xr, pw);
                    pw = fmaf(xi, xi, pw);
                }
            }
        }
        __syncthreads();  // everyone is done with half-block hb-1 and with zb
        if (hb + 1 < nhb) commit(hb + 1);
    }

    // reductions over the 8 warps (alias zb)
    __syncthreads();
    double* dsum = reinterpret_cast<double*>(zb);       // [8][32]
    float* psum = reinterpret_cast<float*>(dsum + 8 * kSFPix);  // [8][32]
    dsum[w * kSFPix + lane] = msum;
    psum[w * kSFPix + lane] = pw;
    __syncthreads();
    if (w == 0 && valid) {
        double m = 0.0;
        float p = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            m += dsum[r * kSFPix + lane];
            p += psum[r * kSFPix + lane];
        }
        const int64_t o = (int64_t)chunk * d + px;
        mean_part[o] = (float)(m * inv_total);
        noise_part[o] = nseg > 0 ? sqrtf(p / (64.f * 96.f * (float)nseg)) : 0.f;
    }
}

}  // namespace pmd

extern "C" int pmd_stats_pass(const void* movie, int dtype, int64_t t_local, int64_t d, int64_t t_total, const float* tab,
                              float* mean_part, float* noise_part, void* stream) {
    const char* fn = "pmd_stats_pass";
    PMD_REQUIRE(movie && tab && mean_part && noise_part, fn, "null pointer");
    PMD_REQUIRE(t_local > 0 && d > 0 && t_total > 0, fn, "non-positive size");
    const int64_t n_chunks = (t_local + pmd::kSFChunk - 1) / pmd::kSFChunk;
    PMD_REQUIRE(n_chunks <= 65535, fn, "too many chunks for one call (t_local > 65535*1024)");
    const size_t smem = (size_t)(256 * pmd::kSFPix + 2 * 128 * pmd::kSFPix + 256 + 256 + 260) * sizeof(float);
    dim3 grid((unsigned)((d + pmd::kSFPix - 1) / pmd::kSFPix), (unsigned)n_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        auto k = pmd::stats_fft_kernel<scalar_t>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
        k<<<grid, pmd::kSFThreads, smem, st>>>((const scalar_t*)movie, t_local, d, 1.0 / (double)t_total, tab, mean_part,
                                                noise_part);
    });
    return pmd::check_launch(fn);
}
