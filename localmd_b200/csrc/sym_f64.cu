// Large SYMMETRIC float64 products on the FP64 tensor cores (mma.sync m8n8k4, DMMA):
//   C[i][j] = sum_k A(i, k) B(j, k),  n x n with n ~ 1650, k_len ~ 1e4 .. 2e4,
// the k x k Gram of the projected movie (decomposition.py:1063-1071, fewer_rows_svd_routine) and the whitening Gram
// M^T (U^T U M) (decomposition.py:974-983).  The operands are float32 (exact in float64) or float64 and are converted
// while they are staged, so no float64 copy of the 1650 x 20000 projection is ever written.  Only the 128 x 128 tiles on
// and above the diagonal are computed (the product is symmetric by construction), the inner dimension is split so that
// the units fill whole waves of 148 SMs, and a second kernel adds the partial tiles in a fixed order (deterministic)
// and writes both triangles.
#include "common.cuh"

namespace pmd {

constexpr int kSymTile = 128;      // output tile edge
constexpr int kSymBK = 16;         // inner-dimension values per shared-memory stage
constexpr int kSymThreads = 256;   // 8 warps as 2 (rows) x 4 (columns); warp tile 64 x 32
constexpr int kSymLd0 = kSymBK + 4;     // layout 0: [row][k], pitch 20 doubles (4 mod 16: the 8 x 4 fragment loads take the
                                        // minimum of two wavefronts)
constexpr int kSymLd1 = kSymTile + 8;   // layout 1: [k][row], pitch 136 doubles (8 mod 16)

template <typename T>
struct Raw8 {
    T v[8];
};

template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ p, int valid, Raw8<T>& r);

template <>
__device__ __forceinline__ void load8<float>(const float* __restrict__ p, int valid, Raw8<float>& r) {
    if (valid >= 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p));
        const float4 y = __ldg(reinterpret_cast<const float4*>(p) + 1);
        r.v[0] = x.x; r.v[1] = x.y; r.v[2] = x.z; r.v[3] = x.w;
        r.v[4] = y.x; r.v[5] = y.y; r.v[6] = y.z; r.v[7] = y.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = e < valid ? __ldg(p + e) : 0.f;
    }
}

template <>
__device__ __forceinline__ void load8<double>(const double* __restrict__ p, int valid, Raw8<double>& r) {
    if (valid >= 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double2 x = __ldg(reinterpret_cast<const double2*>(p) + e);
            r.v[2 * e] = x.x;
            r.v[2 * e + 1] = x.y;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = e < valid ? __ldg(p + e) : 0.0;
    }
}

template <typename T>
__device__ __forceinline__ void store8(double* dst, const Raw8<T>& r) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
        reinterpret_cast<double2*>(dst)[e] = make_double2((double)r.v[2 * e], (double)r.v[2 * e + 1]);
}

// One operand tile (128 rows x 16 inner values) from global memory into registers.
//   LAYOUT 0: element (row, k) at p[row * ld + k]: thread -> row tid / 2, 8 consecutive k
//   LAYOUT 1: element (row, k) at p[k * ld + row]: thread -> k tid / 16, 8 consecutive rows
template <typename T, int LAYOUT>
__device__ __forceinline__ void fetch_tile(const T* __restrict__ p, int64_t ld, int row0, int n, int64_t k0, int64_t k_end,
                                           int tid, Raw8<T>& r) {
    if (LAYOUT == 0) {
        const int row = row0 + (tid >> 1);
        const int64_t k = k0 + (tid & 1) * 8;
        int valid = 0;
        if (row < n && k < k_end) valid = (int)min((int64_t)8, k_end - k);
        load8<T>(p + (int64_t)min(row, n - 1) * ld + min(k, k_end - 1), valid, r);
    } else {
        const int64_t k = k0 + (tid >> 4);
        const int row = row0 + (tid & 15) * 8;
        int valid = 0;
        if (k < k_end && row < n) valid = min(8, n - row);
        load8<T>(p + min(k, k_end - 1) * ld + min(row, n - 1), valid, r);
    }
}

template <typename T, int LAYOUT>
__device__ __forceinline__ void stash_tile(double* sm, int tid, const Raw8<T>& r) {
    if (LAYOUT == 0)
        store8<T>(sm + (tid >> 1) * kSymLd0 + (tid & 1) * 8, r);
    else
        store8<T>(sm + (tid >> 4) * kSymLd1 + (tid & 15) * 8, r);
}

template <int LAYOUT>
__device__ __forceinline__ double frag(const double* sm, int row, int k) {
    return LAYOUT == 0 ? sm[row * kSymLd0 + k] : sm[k * kSymLd1 + row];
}

template <typename TA, typename TB, int LAYOUT>
__global__ void __launch_bounds__(kSymThreads, 1)
sym_product_dmma_kernel(const TA* __restrict__ a, int64_t lda, const TB* __restrict__ b, int64_t ldb, int n, int64_t k_len,
                        int64_t k_chunk, int nt, int ntiles, double* __restrict__ out) {
    extern __shared__ double ssm[];
    constexpr int kOpDoubles = LAYOUT == 0 ? kSymTile * kSymLd0 : kSymBK * kSymLd1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
    // unit = (split, tile): consecutive CTAs work on the SAME inner-dimension chunk, which L2 then serves to all tiles
    const int split = blockIdx.x / ntiles;
    int tile = blockIdx.x - split * ntiles;
    int ti = 0;
    while (tile >= nt - ti) { tile -= nt - ti; ++ti; }
    const int tj = ti + tile;
    const int64_t k_begin = (int64_t)split * k_chunk;
    const int64_t k_end = min(k_len, k_begin + k_chunk);
    const int nk = (int)((k_end - k_begin + kSymBK - 1) / kSymBK);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    Raw8<TA> ra;
    Raw8<TB> rb;
    if (nk > 0) {
        fetch_tile<TA, LAYOUT>(a, lda, ti * kSymTile, n, k_begin, k_end, tid, ra);
        fetch_tile<TB, LAYOUT>(b, ldb, tj * kSymTile, n, k_begin, k_end, tid, rb);
        stash_tile<TA, LAYOUT>(ssm, tid, ra);
        stash_tile<TB, LAYOUT>(ssm + kOpDoubles, tid, rb);
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const double* sa = ssm + (kt & 1) * 2 * kOpDoubles;
        const double* sb = sa + kOpDoubles;
        if (kt + 1 < nk) {
            const int64_t k0 = k_begin + (int64_t)(kt + 1) * kSymBK;
            fetch_tile<TA, LAYOUT>(a, lda, ti * kSymTile, n, k0, k_end, tid, ra);
            fetch_tile<TB, LAYOUT>(b, ldb, tj * kSymTile, n, k0, k_end, tid, rb);
        }
#pragma unroll
        for (int kk = 0; kk < kSymBK; kk += 4) {
            double fa[8], fb[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) fa[i] = frag<LAYOUT>(sa, wm + 8 * i + fr, kk + fc);
#pragma unroll
            for (int j = 0; j < 4; ++j) fb[j] = frag<LAYOUT>(sb, wn + 8 * j + fr, kk + fc);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                                 : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                                 : "d"(fa[i]), "d"(fb[j]));
        }
        if (kt + 1 < nk) {
            double* da = ssm + ((kt + 1) & 1) * 2 * kOpDoubles;
            stash_tile<TA, LAYOUT>(da, tid, ra);
            stash_tile<TB, LAYOUT>(da + kOpDoubles, tid, rb);
        }
        __syncthreads();
    }
    // partial tile [128][128] of this unit; lane holds C[fr][2 fc], C[fr][2 fc + 1] of every 8 x 8 fragment
    double* o = out + (int64_t)blockIdx.x * kSymTile * kSymTile;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(o + (wm + 8 * i + fr) * kSymTile + wn + 8 * j + 2 * fc) =
                make_double2(acc[i][j][0], acc[i][j][1]);
}

// C[i][j] = C[j][i] = sum over the splits (ascending) of the partial tiles, for the elements with j >= i
__global__ void __launch_bounds__(256)
sym_reduce_kernel(const double* __restrict__ part, int n, int nt, int ntiles, int splits, double* __restrict__ c) {
    int tile = blockIdx.x;
    int ti = 0;
    while (tile >= nt - ti) { tile -= nt - ti; ++ti; }
    const int tj = ti + tile;
    const double* p = part + (int64_t)blockIdx.x * kSymTile * kSymTile;
    const int64_t split_stride = (int64_t)ntiles * kSymTile * kSymTile;
    for (int e = threadIdx.x; e < kSymTile * kSymTile; e += blockDim.x) {
        const int li = e >> 7, lj = e & 127;
        const int i = ti * kSymTile + li, j = tj * kSymTile + lj;
        if (i >= n || j >= n || j < i) continue;
        double s = 0.0;
        for (int q = 0; q < splits; ++q) s += p[q * split_stride + e];
        c[(int64_t)i * n + j] = s;
        if (i != j) c[(int64_t)j * n + i] = s;
    }
}

template <typename TA, typename TB, int LAYOUT>
static int launch_sym(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t n, int64_t k_len, int64_t splits,
                      double* work, double* c, cudaStream_t stream, const char* fn) {
    const int nt = (int)((n + kSymTile - 1) / kSymTile);
    const int ntiles = nt * (nt + 1) / 2;
    int64_t k_chunk = (k_len + splits - 1) / splits;
    k_chunk = (k_chunk + kSymBK - 1) / kSymBK * kSymBK;
    constexpr size_t smem = sizeof(double) * 4 * (LAYOUT == 0 ? kSymTile * kSymLd0 : kSymBK * kSymLd1);
    auto kern = sym_product_dmma_kernel<TA, TB, LAYOUT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error(std::string(fn) + ": " + cudaGetErrorString(e));
        return (int)e;
    }
    kern<<<(unsigned)(ntiles * splits), kSymThreads, smem, stream>>>((const TA*)a, lda, (const TB*)b, ldb, (int)n, k_len, k_chunk,
                                                                      nt, ntiles, work);
    int rc = check_launch(fn);
    if (rc) return rc;
    sym_reduce_kernel<<<ntiles, 256, 0, stream>>>(work, (int)n, nt, ntiles, (int)splits, c);
    return check_launch(fn);
}

}  // namespace pmd

extern "C" int pmd_sym_product_f64(const void* a, int a_dtype, int64_t lda, const void* b, int b_dtype, int64_t ldb, int layout,
                                   int64_t n, int64_t k_len, int64_t splits, double* work, double* c, void* stream) {
    const char* fn = "pmd_sym_product_f64";
    PMD_REQUIRE(a && work && c, fn, "null pointer");
    PMD_REQUIRE(n >= 1 && n <= 128 * 255 && k_len >= 1, fn, "n in 1..32640, k_len >= 1");
    PMD_REQUIRE(splits >= 1 && splits <= 64, fn, "splits in 1..64");
    PMD_REQUIRE(layout == 0 || layout == 1, fn, "layout 0 (inner dimension contiguous) or 1 (rows contiguous)");
    if (!b) {
        b = a;
        b_dtype = a_dtype;
        ldb = lda;
    }
    PMD_REQUIRE(layout == 0 ? (lda >= k_len && ldb >= k_len) : (lda >= n && ldb >= n), fn, "leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (a_dtype == PMD_F32 && b_dtype == PMD_F32 && layout == 0)
        return pmd::launch_sym<float, float, 0>(a, lda, b, ldb, n, k_len, splits, work, c, st, fn);
    if (a_dtype == PMD_F64 && b_dtype == PMD_F64 && layout == 0)
        return pmd::launch_sym<double, double, 0>(a, lda, b, ldb, n, k_len, splits, work, c, st, fn);
    if (a_dtype == PMD_F32 && b_dtype == PMD_F64 && layout == 1)
        return pmd::launch_sym<float, double, 1>(a, lda, b, ldb, n, k_len, splits, work, c, st, fn);
    if (a_dtype == PMD_F64 && b_dtype == PMD_F64 && layout == 1)
        return pmd::launch_sym<double, double, 1>(a, lda, b, ldb, n, k_len, splits, work, c, st, fn);
    return pmd::fail_arg(fn, "supported operand types: layout 0 f32/f32, f64/f64; layout 1 f32/f64, f64/f64");
}
