// K1: per-pixel mean + Welch high-band noise estimate in one streaming pass over the movie.
//
// Replaces pmd_loader.py:203-291 and preprocessing_utils.py:10-40 of the reference.
//
// Per 1024-frame chunk and pixel the reference computes scipy/jax `welch(trace, noverlap=128)`
// (Hann-periodic window, 256-sample segments, hop 128, constant detrend, one-sided density, mean
// over segments) and then averages 0.5*Pxx over bins 65..128.  Two identities make this a small
// dense contraction per segment instead of an FFT:
//   * the Hann-windowed DFT of a constant is non-zero only at bins 0 and +-1, so the per-segment
//     mean removal does not change bins >= 65 (we subtract a per-pixel offset only for rounding);
//   * w[n] = w[256-n] and cos/sin are even/odd about n = 128, so with e[n] = x[n] + x[256-n],
//     o[n] = x[n] - x[256-n] (n = 1..127), e[128] = x[128]:
//         Re X[k] = sum_{n=1..128} (w[n] cos(2 pi k n/256)) e[n],   |Im X[k]| = |sum (w[n] sin(..)) o[n]|
//     i.e. two [64 bins] x [128] tables applied to every pixel's folded segment.
// sigma^2 = 1/(64*96*nseg) * sum_seg ( sum_{k=65..127} |X[k]|^2 + 0.5 |X[128]|^2 )   (96 = sum w^2).
//
// One CTA = 64 pixels x one chunk.  256 threads; the 256-frame ring buffer, both tables and the
// reductions live in shared memory (132 KB).  Threads 0..127 evaluate the cosine table, 128..255
// the sine table, each as an 8(bins) x 4(pixels) register tile.
#include "common.cuh"

namespace pmd {

constexpr int kStatsPix = 64;
constexpr int kStatsThreads = 256;
constexpr int kChunk = 1024;
constexpr int kSeg = 256;
constexpr int kHop = 128;

template <typename T>
__global__ void __launch_bounds__(kStatsThreads, 1)
stats_kernel(const T* __restrict__ movie, int64_t t_local, int64_t d, double inv_total,
             const float* __restrict__ tab_cos, const float* __restrict__ tab_sin,
             float* __restrict__ mean_part, float* __restrict__ noise_part) {
    extern __shared__ __align__(16) float smem[];
    float* tabc = smem;                   // [128][64]
    float* tabs = smem + 128 * 64;        // [128][64]
    float* ring = smem + 2 * 128 * 64;    // [256][64]

    const int tid = threadIdx.x;
    const int64_t px0 = (int64_t)blockIdx.x * kStatsPix;
    const int chunk = blockIdx.y;
    const int64_t f_begin = (int64_t)chunk * kChunk;
    const int n = (int)min((int64_t)kChunk, t_local - f_begin);
    const int nseg = n >= kSeg ? (n - kHop) / kHop : 0;

    for (int i = tid; i < 128 * 64; i += kStatsThreads) {
        tabc[i] = tab_cos[i];
        tabs[i] = tab_sin[i];
    }

    // loader mapping: 16 pixel quads x 16 frame rows
    const int lq = tid & 15, lr = tid >> 4;
    const int64_t lpx = px0 + lq * 4;
    float c0[4];
    bool valid[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        valid[j] = (lpx + j) < d;
        c0[j] = valid[j] ? to_f32(movie[f_begin * d + lpx + j]) : 0.f;
    }
    double msum[4] = {0.0, 0.0, 0.0, 0.0};

    // compute mapping
    const int grp = tid >> 7;  // 0 cosine table, 1 sine table
    const int g = tid & 127;
    const int pq = g & 15, rg = g >> 4;
    const float* tab = grp ? tabs : tabc;
    float pw[4] = {0.f, 0.f, 0.f, 0.f};

    const int nhb = (n + kHop - 1) / kHop;
    for (int hb = 0; hb < nhb; ++hb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int fr = hb * kHop + lr + 16 * i;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (fr < n) {
                const T* src = movie + (f_begin + fr) * d + lpx;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (valid[j]) {
                        const float x = to_f32(src[j]);
                        msum[j] += (double)x;
                        v[j] = x - c0[j];
                    }
                }
            }
            *reinterpret_cast<float4*>(&ring[(fr & 255) * 64 + lq * 4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        const int s = hb - 1;
        if (s >= 0 && s < nseg) {
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            const int base = s * kHop;
#pragma unroll 4
            for (int nn = 1; nn <= 128; ++nn) {
                const float4 a = *reinterpret_cast<const float4*>(&ring[((base + nn) & 255) * 64 + pq * 4]);
                float4 b = *reinterpret_cast<const float4*>(&ring[((base + 256 - nn) & 255) * 64 + pq * 4]);
                float x[4];
                if (grp == 0) {
                    if (nn == 128) b = make_float4(0.f, 0.f, 0.f, 0.f);
                    x[0] = a.x + b.x; x[1] = a.y + b.y; x[2] = a.z + b.z; x[3] = a.w + b.w;
                } else {
                    x[0] = a.x - b.x; x[1] = a.y - b.y; x[2] = a.z - b.z; x[3] = a.w - b.w;
                }
                const float4 t0 = *reinterpret_cast<const float4*>(&tab[(nn - 1) * 64 + rg * 8]);
                const float4 t1 = *reinterpret_cast<const float4*>(&tab[(nn - 1) * 64 + rg * 8 + 4]);
                const float tv[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(tv[i], x[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float coef = (grp == 0 && rg == 7 && i == 7) ? 0.5f : 1.f;  // Nyquist bin is not doubled
#pragma unroll
                for (int j = 0; j < 4; ++j) pw[j] = fmaf(coef * acc[i][j], acc[i][j], pw[j]);
            }
        }
        __syncthreads();
    }

    // reductions (alias the ring buffer: everyone is past the last barrier)
    double* dsum = reinterpret_cast<double*>(ring);           // [16][64]
    float* psum = ring + 2 * 16 * 64;                         // [16][64]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        dsum[lr * 64 + lq * 4 + j] = msum[j];
        psum[(grp * 8 + rg) * 64 + pq * 4 + j] = pw[j];
    }
    __syncthreads();
    if (tid < kStatsPix && px0 + tid < d) {
        double m = 0.0;
        float p = 0.f;
        for (int r = 0; r < 16; ++r) {
            m += dsum[r * 64 + tid];
            p += psum[r * 64 + tid];
        }
        const int64_t o = (int64_t)chunk * d + px0 + tid;
        mean_part[o] = (float)(m * inv_total);
        noise_part[o] = nseg > 0 ? sqrtf(p / (64.f * 96.f * (float)nseg)) : 0.f;
    }
}

template <typename T>
__global__ void standardize_frames_kernel(const T* __restrict__ movie, int64_t d, const int64_t* __restrict__ frames,
                                          const float* __restrict__ mean, const float* __restrict__ stdv,
                                          float* __restrict__ out) {
    const int64_t i = blockIdx.y;
    const int64_t f = frames[i];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < d; p += (int64_t)gridDim.x * blockDim.x) {
        out[i * d + p] = (to_f32(movie[f * d + p]) - mean[p]) / stdv[p];
    }
}

}  // namespace pmd

extern "C" int pmd_stats_pass(const void* movie, int dtype, int64_t t_local, int64_t d, int64_t t_total,
                              const float* tab_cos, const float* tab_sin, float* mean_part, float* noise_part,
                              void* stream) {
    const char* fn = "pmd_stats_pass";
    PMD_REQUIRE(movie && tab_cos && tab_sin && mean_part && noise_part, fn, "null pointer");
    PMD_REQUIRE(t_local > 0 && d > 0 && t_total > 0, fn, "non-positive size");
    const int64_t n_chunks = (t_local + pmd::kChunk - 1) / pmd::kChunk;
    PMD_REQUIRE(n_chunks <= 65535, fn, "too many chunks for one call (t_local > 65535*1024)");
    const size_t smem = (size_t)(2 * 128 * 64 + 256 * 64) * sizeof(float);
    dim3 grid((unsigned)((d + pmd::kStatsPix - 1) / pmd::kStatsPix), (unsigned)n_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        auto k = pmd::stats_kernel<scalar_t>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
        k<<<grid, pmd::kStatsThreads, smem, st>>>((const scalar_t*)movie, t_local, d, 1.0 / (double)t_total, tab_cos,
                                                   tab_sin, mean_part, noise_part);
    });
    return pmd::check_launch(fn);
}

extern "C" int pmd_standardize_frames(const void* movie, int dtype, int64_t d, const int64_t* frames, int64_t n_frames,
                                      const float* mean, const float* stdv, float* out, void* stream) {
    const char* fn = "pmd_standardize_frames";
    PMD_REQUIRE(movie && frames && mean && stdv && out, fn, "null pointer");
    PMD_REQUIRE(d > 0 && n_frames >= 0, fn, "bad size");
    if (n_frames == 0) return 0;
    PMD_REQUIRE(n_frames <= 65535, fn, "more than 65535 frames per call");
    dim3 grid((unsigned)std::min<int64_t>((d + 255) / 256, 1024), (unsigned)n_frames);
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        pmd::standardize_frames_kernel<scalar_t><<<grid, 256, 0, st>>>((const scalar_t*)movie, d, frames, mean, stdv, out);
    });
    return pmd::check_launch(fn);
}
