// Background basis (pmd_loader.py:46-68: rSVD of <= 1000 standardised frames) on the pixel-major layout of the init
// movie: the sampled frames are standardised + transposed once (pmd_standardize_frames_t), after which both skinny
// contractions of the randomised SVD stream that (d, ld) matrix exactly once each:
//     y[p][j]    = sum_f  yt[p][f] * omega[f][j]              (sketch, l = K + 10 <= 32 columns)           pmd_rows_sketch
//     b[j][f]    = sum_p  q[j][p]  * yt[p][f]                 (coefficients of the orthonormalised sketch)  pmd_bg_project_t
// plus the small right-multiplication  out = x @ m  (m at most 32 x 32) used by the orthonormalisation passes and the
// final rotation into singular vectors                                                                     pmd_rows_times_small
// All are bound by one pass over yt / x; the library GEMMs picked for these (d x 1000) x (1000 x 25) shapes take 2-3x the
// streaming time and need the frame-major standardised copy besides.
#include "common.cuh"

namespace pmd {

constexpr int kRSThreads = 256;      // 8 warps, four pixel rows per warp in flight
constexpr int kRSL = 32;             // max sketch columns
constexpr int kRSChunk = 1024;       // frames of omega resident in shared memory (the whole sketch for <= 1024 frames)
constexpr int kRSPitch = 36;         // row pitch in floats: 128-bit reads of 8 consecutive rows cover all 32 banks
constexpr int kRSRows = 4;           // pixel rows accumulated at once by a warp (one 128-bit omega read feeds 16 FMAs)

// One warp per group of four pixel rows: lane g reads frames 4g .. 4g+3 (+128 per step) of its rows as one 16-byte load (512-byte
// requests per warp), multiplies them with the matching rows of omega (128-bit conflict-free shared-memory reads) and keeps
// 4 * L4 partial sums per row; the warp adds the partial sums by shuffles.  omega is staged once per CTA in the order
// [f % 4][f / 4][kRSPitch].
__device__ __forceinline__ uint64_t rs_pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void rs_unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t rs_fma2(uint64_t a, uint64_t b, uint64_t c) {   // packed float32 pair FMA (Blackwell)
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int L4>
__global__ void __launch_bounds__(kRSThreads)
rows_sketch_kernel(const float* __restrict__ yt, int64_t ld, int64_t d, int64_t n, const float* __restrict__ omega, int l,
                   float* __restrict__ y, int64_t ldy, int64_t rows_per_cta) {
    extern __shared__ __align__(16) float so[];
    constexpr int L = 4 * L4;
    constexpr int NV = kRSRows * L;          // partial sums per lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t p0 = (int64_t)blockIdx.x * rows_per_cta, p1 = min(d, p0 + rows_per_cta);
    for (int64_t f0 = 0; f0 < n; f0 += kRSChunk) {
        const int nf = (int)min((int64_t)kRSChunk, n - f0);
        __syncthreads();
        for (int i = threadIdx.x; i < kRSChunk * L; i += kRSThreads) {
            const int f = i / L, j = i - f * L;
            so[((f & 3) * (kRSChunk / 4) + (f >> 2)) * kRSPitch + j] = (f < nf && j < l) ? omega[(f0 + f) * l + j] : 0.f;
        }
        __syncthreads();
        const int steps = (nf + 127) / 128;
        // (warp-uniform loop: every lane of a warp runs the same trip count, so the shuffles below are convergent)
        for (int64_t pb = p0 + kRSRows * warp; pb < p1; pb += kRSRows * (kRSThreads / 32)) {
            uint64_t acc[kRSRows][L / 2];    // packed pairs (column 2c, 2c + 1)
#pragma unroll
            for (int r = 0; r < kRSRows; ++r)
#pragma unroll
                for (int j = 0; j < L / 2; ++j) acc[r][j] = 0ull;
            auto load = [&](float4 (&v)[kRSRows], int s) {
                const int64_t f = f0 + 4 * (s * 32 + lane);
#pragma unroll
                for (int r = 0; r < kRSRows; ++r)
                    v[r] = (pb + r < p1 && s < steps && f < ld) ? __ldg(reinterpret_cast<const float4*>(yt + (pb + r) * ld + f))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                if (f + 3 >= n) {   // padding columns of yt may hold anything
#pragma unroll
                    for (int r = 0; r < kRSRows; ++r) {
                        if (f + 0 >= n) v[r].x = 0.f;
                        if (f + 1 >= n) v[r].y = 0.f;
                        if (f + 2 >= n) v[r].z = 0.f;
                        v[r].w = 0.f;
                    }
                }
            };
            float4 cur[kRSRows], nxt[kRSRows];
            load(cur, 0);
            for (int s = 0; s < steps; ++s) {
                load(nxt, s + 1);
                const int g = s * 32 + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const ulonglong2* o = reinterpret_cast<const ulonglong2*>(so + (q * (kRSChunk / 4) + g) * kRSPitch);
                    uint64_t vv[kRSRows];
#pragma unroll
                    for (int r = 0; r < kRSRows; ++r) {
                        const float v = q == 0 ? cur[r].x : q == 1 ? cur[r].y : q == 2 ? cur[r].z : cur[r].w;
                        vv[r] = rs_pack2(v, v);
                    }
#pragma unroll
                    for (int j4 = 0; j4 < L4; ++j4) {
                        const ulonglong2 w = o[j4];      // omega[f][4 j4 .. 4 j4 + 3] as two packed pairs
#pragma unroll
                        for (int r = 0; r < kRSRows; ++r) {
                            acc[r][2 * j4 + 0] = rs_fma2(vv[r], w.x, acc[r][2 * j4 + 0]);
                            acc[r][2 * j4 + 1] = rs_fma2(vv[r], w.y, acc[r][2 * j4 + 1]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < kRSRows; ++r) cur[r] = nxt[r];
            }
            // reduce-scatter over the warp: after 5 halving exchanges lane b holds the totals of NV / 32 consecutive values
            // of the flattened [row][column] array (padded to a multiple of 32), starting at its bit-reversed-free index
            constexpr int NP = (NV + 31) / 32 * 32;
            float v[NP];
#pragma unroll
            for (int r = 0; r < kRSRows; ++r)
#pragma unroll
                for (int j = 0; j < L / 2; ++j) rs_unpack2(acc[r][j], v[r * L + 2 * j], v[r * L + 2 * j + 1]);
#pragma unroll
            for (int i = NV; i < NP; ++i) v[i] = 0.f;
            int base = 0;
#pragma unroll
            for (int o = 16, cnt = NP / 2; o >= 1; o >>= 1, cnt >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < cnt; ++i) {
                    const float send = up ? v[i] : v[i + cnt];
                    const float keep = up ? v[i + cnt] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
                base += up ? cnt : 0;
            }
#pragma unroll
            for (int i = 0; i < NP / 32; ++i) {
                const int idx = base + i, r = idx / L, j = idx - r * L;
                if (idx < NV && pb + r < p1 && j < l) {
                    float* dst = y + (pb + r) * ldy + j;
                    *dst = f0 == 0 ? v[i] : *dst + v[i];
                }
            }
        }
    }
}

// out[p][c] = sum_j x[p][j] m[j][c]   (j < k <= KM, c < nc <= KM; KM = 32); out may be written transposed
// (out_t[c][p]).  One thread per row: the row of x lives in registers, m in shared memory (broadcast reads).
constexpr int kRTThreads = 128;

template <int KM>
__global__ void __launch_bounds__(kRTThreads)
rows_times_small_kernel(const float* __restrict__ x, int64_t ldx, int64_t d, int k, const float* __restrict__ m, int64_t ldm, int nc,
                        float* __restrict__ out, int64_t ldo, int transposed, int64_t batch_stride_x, int64_t batch_stride_m,
                        int64_t batch_stride_o) {
    __shared__ float sm[KM * KM];
    const int64_t b = blockIdx.y;
    x += b * batch_stride_x;
    m += b * batch_stride_m;
    out += b * batch_stride_o;
    for (int i = threadIdx.x; i < KM * KM; i += kRTThreads) {
        const int j = i / KM, c = i - j * KM;
        sm[i] = (j < k && c < nc) ? m[(int64_t)j * ldm + c] : 0.f;
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * kRTThreads + threadIdx.x;
    if (p >= d) return;
    float xr[KM];
#pragma unroll
    for (int j = 0; j < KM; ++j) xr[j] = j < k ? x[p * ldx + j] : 0.f;
    for (int c = 0; c < nc; ++c) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < KM; ++j) s = fmaf(xr[j], sm[j * KM + c], s);
        if (transposed) out[(int64_t)c * ldo + p] = s;
        else out[p * ldo + c] = s;
    }
}

// t = L^-T of the Cholesky factor g = L L^T of a small float64 Gram matrix (n <= 32): x @ t has orthonormal columns when
// g = x^T x.  One warp per matrix, the matrix in shared memory.  A column whose pivot falls below 1e-13 of the largest
// diagonal entry (numerically dependent on the earlier ones) gives a zero column of t.
__global__ void __launch_bounds__(32)
chol_whiten_kernel(const double* __restrict__ g, int n, float* __restrict__ t) {
    __shared__ double a[kRSL][kRSL + 1], x[kRSL][kRSL + 1];
    __shared__ int dead[kRSL];
    const int lane = threadIdx.x;
    g += (int64_t)blockIdx.x * n * n;
    t += (int64_t)blockIdx.x * n * n;
    double dmax = 0.0;
    for (int j = 0; j < n; ++j) {
        if (lane < n) a[lane][j] = g[(int64_t)lane * n + j];
        dmax = fmax(dmax, g[(int64_t)j * n + j]);
    }
    __syncwarp();
    for (int j = 0; j < n; ++j) {
        const double piv = a[j][j];
        const bool bad = !(piv > 1e-13 * dmax);
        const double dj = bad ? 1.0 : sqrt(piv);
        __syncwarp();
        if (lane == j) { a[j][j] = dj; dead[j] = bad; }
        if (lane > j && lane < n) a[lane][j] = bad ? 0.0 : a[lane][j] / dj;
        __syncwarp();
        if (lane > j && lane < n) {
            const double lij = a[lane][j];
            for (int k = j + 1; k <= lane; ++k) a[lane][k] -= lij * a[k][j];
        }
        __syncwarp();
    }
    // column c of X = L^-1 by forward substitution (lane c)
    if (lane < n) {
        for (int i = 0; i < n; ++i) {
            double sum = i == lane ? 1.0 : 0.0;
            for (int k = lane; k < i; ++k) sum -= a[i][k] * x[k][lane];
            x[i][lane] = i < lane ? 0.0 : sum / a[i][i];
        }
    }
    __syncwarp();
    // t[j][c] = X[c][j]; dead columns c are zeroed
    for (int j = 0; j < n; ++j)
        if (lane < n) t[(int64_t)j * n + lane] = dead[lane] ? 0.f : (float)x[lane][j];
}

}  // namespace pmd

extern "C" int pmd_chol_whiten(const double* g, int64_t batch, int64_t n, float* t, void* stream) {
    const char* fn = "pmd_chol_whiten";
    PMD_REQUIRE(g && t, fn, "null pointer");
    PMD_REQUIRE(batch > 0 && n > 0 && n <= pmd::kRSL, fn, "bad size (1 <= n <= 32)");
    pmd::chol_whiten_kernel<<<(unsigned)batch, 32, 0, (cudaStream_t)stream>>>(g, (int)n, t);
    return pmd::check_launch(fn);
}

extern "C" int pmd_rows_sketch(const float* yt, int64_t ld, int64_t d, int64_t n, const float* omega, int64_t l, float* y, int64_t ldy,
                               void* stream) {
    const char* fn = "pmd_rows_sketch";
    PMD_REQUIRE(yt && omega && y, fn, "null pointer");
    PMD_REQUIRE(d > 0 && n > 0 && n <= ld && l > 0 && l <= pmd::kRSL && ldy >= l, fn, "bad size (1 <= l <= 32, n <= ld, ldy >= l)");
    PMD_REQUIRE(ld % 4 == 0 && ((uintptr_t)yt % 16) == 0, fn, "yt rows must be 16-byte aligned (ld a multiple of 4)");
    // one persistent CTA per SM (omega resident in its shared memory), row ranges in multiples of 32 rows
    int64_t per = (d + 148 * 2 - 1) / (148 * 2);
    per = (per + 31) / 32 * 32;
    const unsigned grid = (unsigned)((d + per - 1) / per);
    const int smem = pmd::kRSChunk * pmd::kRSPitch * (int)sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
#define PMD_RS_LAUNCH(L4)                                                                                                   \
    do {                                                                                                                     \
        cudaError_t e = cudaFuncSetAttribute(pmd::rows_sketch_kernel<L4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) {                                                                                              \
            pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e));                                                  \
            return (int)e;                                                                                                   \
        }                                                                                                                    \
        pmd::rows_sketch_kernel<L4><<<grid, pmd::kRSThreads, smem, st>>>(yt, ld, d, n, omega, (int)l, y, ldy, per);          \
    } while (0)
    if (l <= 16) PMD_RS_LAUNCH(4);
    else if (l <= 28) PMD_RS_LAUNCH(7);   // the default sketch width: background_rank 15 + 10 = 25 columns
    else PMD_RS_LAUNCH(8);
#undef PMD_RS_LAUNCH
    return pmd::check_launch(fn);
}

extern "C" int pmd_rows_times_small(const float* x, int64_t ldx, int64_t d, int64_t k, const float* m, int64_t ldm, int64_t nc,
                                    float* out, int64_t ldo, int transposed, int64_t batch, int64_t batch_stride_x,
                                    int64_t batch_stride_m, int64_t batch_stride_o, void* stream) {
    const char* fn = "pmd_rows_times_small";
    PMD_REQUIRE(x && m && out, fn, "null pointer");
    PMD_REQUIRE(d > 0 && k > 0 && k <= 32 && nc > 0 && nc <= 32 && ldx >= k && ldm >= nc, fn, "bad size (k, nc <= 32)");
    PMD_REQUIRE(batch > 0 && batch <= 65535, fn, "batch must be in 1..65535");
    PMD_REQUIRE(x != out, fn, "in-place operation is not supported");
    dim3 grid((unsigned)((d + pmd::kRTThreads - 1) / pmd::kRTThreads), (unsigned)batch);
    pmd::rows_times_small_kernel<32><<<grid, pmd::kRTThreads, 0, (cudaStream_t)stream>>>(
        x, ldx, d, (int)k, m, ldm, (int)nc, out, ldo, transposed, batch_stride_x, batch_stride_m, batch_stride_o);
    return pmd::check_launch(fn);
}
