// Batched orthonormalisation of tall-skinny matrices that fit in shared memory (m x n, n <= 64,
// m * n * 4 bytes <= ~150 KB), one CTA per matrix:
//   (optional) X <- X L_g^-T          with  G_ext = L_g L_g^T  given in float64
//   twice:     X <- X L^-T            with  X^T X = L L^T       (CholQR2; Gram in float64)
// replaces the QR factorisations / SVD-based bases of decomposition.py:64 (Q of the sketch), 301 (temporal
// basis: only its row space is used downstream) and 315 (spatial basis), where any orthonormal basis of the
// same column space gives the same U, V: the final per-block SVD (decomposition.py:318-323) re-diagonalises.
// Products of float32 numbers are exact in float64, so the Gram matrix carries only summation rounding
// (1e-16); a column whose residual after projecting out its predecessors is below float32 noise
// (pivot <= 1e-13 * its squared norm) is DROPPED (set to zero) instead of being normalised -- the behaviour
// of a rank-revealing QR on exactly rank-deficient inputs.
#include "common.cuh"

// Both contractions run on the FP64 tensor cores (mma.sync m8n8k4, float32 inputs converted on the fragment load: exact):
//   Gram          G = X^T X         upper 8 x 8 tiles shared out over the 8 warps, two independent accumulator chains
//   solve         X <- X T^T        with T = L^-1 formed explicitly in the unused upper triangle of G (four lanes per
//                                   column); a warp owns 8 rows of X, keeps them in registers as A fragments and writes
//                                   the product back in place
// (the first version -- one thread per Gram entry / per row of the substitution, scalar float64 FMAs fed by two shared
// memory loads each -- issued 685 k warp instructions per 400 x 50 matrix; ncu: 3.4 ms for the 2601 blocks of C2).
namespace pmd {

constexpr int kOrthThreads = 256;

__device__ __forceinline__ void orth_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// in-place Cholesky of the n x n float64 matrix g (row-major, ld), lower triangle (the diagonal of g is left as the
// pivots; 1 / L_jj goes to dinv).  A column whose pivot is <= tol * diag0[j] is dropped: its column of L is zero and
// dinv[j] = 0.  colj: scratch [n].
__device__ void chol_inplace(double* g, int n, int ld, double* dinv, const double* diag0, double* colj, double tol) {
    const int tid = threadIdx.x;
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        const double piv = g[j * ld + j];
        const bool ok = piv > tol * diag0[j] && piv > 0.0;
        const double inv = ok ? rsqrt(piv) : 0.0;
        if (tid == 0) dinv[j] = inv;
        for (int i = j + 1 + tid; i < n; i += kOrthThreads) {
            const double v = g[i * ld + j] * inv;
            g[i * ld + j] = v;
            colj[i] = v;
        }
        __syncthreads();
        // trailing update of the lower triangle: g[a][b] -= l[a][j] * l[b][j], j < b <= a < n
        for (int a = j + 1 + (tid >> 4); a < n; a += kOrthThreads >> 4) {
            const double la = colj[a];
            for (int b = j + 1 + (tid & 15); b <= a; b += 16) g[a * ld + b] -= la * colj[b];
        }
    }
    __syncthreads();
}

// T = L^-1 (with 1 / L_jj := dinv[j], 0 for dropped columns) into the strict UPPER triangle of g: T[i][c], i > c, at
// g[c][i]; T[c][c] = dinv[c] stays in dinv.  Four lanes per column split the inner sums.
__device__ void tri_inverse_upper(double* g, int n, int ld, const double* dinv) {
    const int c = threadIdx.x >> 2, sub = threadIdx.x & 3;
    for (int i = 1; i < n; ++i) {
        double part = 0.0;
        if (c < i) {
            for (int k = c + sub; k < i; k += 4) part += g[i * ld + k] * (k == c ? dinv[c] : g[c * ld + k]);
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (c < i && sub == 0) g[c * ld + i] = -dinv[i] * part;
        __syncwarp();
    }
    __syncthreads();
}

// X <- X T^T for the rows of X in shared memory ([m8][lds] float32), T as left by tri_inverse_upper
__device__ void apply_inverse(float* xs, int m8, int n4, int n8, int lds, const double* g, int ld, const double* dinv) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int nkt = n4 >> 2;
    for (int rt = warp; rt < (m8 >> 3); rt += kOrthThreads >> 5) {
        float* xr = xs + (size_t)(8 * rt + fr) * lds;
        double a[16];
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) a[ks] = ks < nkt ? (double)xr[4 * ks + fc] : 0.0;
        __syncwarp();
        for (int ct = 0; ct < (n8 >> 3); ++ct) {
            const int j = 8 * ct + fr;            // output column of this lane's B fragment element
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                if (ks < nkt && 4 * ks <= 8 * ct + 7) {        // T[j][k] = 0 for k > j
                    const int k = 4 * ks + fc;
                    const double b = k < j ? g[k * ld + j] : (k == j ? dinv[j] : 0.0);
                    orth_dmma(c0, c1, a[ks], b);
                }
            }
            const int col = 8 * ct + 2 * fc;
            if (col < lds) *reinterpret_cast<float2*>(xr + col) = make_float2((float)c0, (float)c1);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kOrthThreads, 2)
block_orth_kernel(float* __restrict__ x, int m, int n, int ldx, const double* __restrict__ g_ext, int passes) {
    extern __shared__ __align__(16) unsigned char osm[];
    const int n4 = (n + 3) & ~3, n8 = (n + 7) & ~7, m8 = (m + 7) & ~7;
    const int ldg = (n8 & 8) ? n8 : n8 + 8;             // 8 mod 16 doubles: two-wavefront fragment loads of T
    const int lds = n4;
    double* g = reinterpret_cast<double*>(osm);          // [n8][ldg]
    double* dinv = g + (size_t)n8 * ldg;                 // [n8]
    double* diag0 = dinv + n8;                           // [n8]
    double* colj = diag0 + n8;                           // [n8]
    float* xs = reinterpret_cast<float*>(colj + n8);     // [m8][lds]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    float* xb = x + (size_t)blockIdx.x * m * ldx;

    for (int idx = tid; idx < n8 * ldg; idx += kOrthThreads) g[idx] = 0.0;
    if (tid < n8) dinv[tid] = 0.0;
    const bool vec = (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0;
    const int nq = lds >> 2;
    for (int idx = tid; idx < m8 * nq; idx += kOrthThreads) {
        const int r = idx / nq, q = idx - r * nq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < m) {
            const float* src = xb + (size_t)r * ldx + 4 * q;
            if (vec && 4 * q + 3 < ldx) {
                v = *reinterpret_cast<const float4*>(src);
            } else {
                if (4 * q < n) v.x = src[0];
                if (4 * q + 1 < n) v.y = src[1];
                if (4 * q + 2 < n) v.z = src[2];
                if (4 * q + 3 < n) v.w = src[3];
            }
            if (4 * q + 1 >= n) v.y = 0.f;               // the padding columns n .. n4 - 1 take no part
            if (4 * q + 2 >= n) v.z = 0.f;
            if (4 * q + 3 >= n) v.w = 0.f;
        }
        *reinterpret_cast<float4*>(xs + (size_t)r * lds + 4 * q) = v;
    }
    __syncthreads();
    if (g_ext) {
        const double* ge = g_ext + (size_t)blockIdx.x * n * n;
        for (int idx = tid; idx < n * n; idx += kOrthThreads) {
            const int i = idx / n, j = idx - i * n;
            if (j <= i) g[i * ldg + j] = ge[idx];
        }
        __syncthreads();
        if (tid < n) diag0[tid] = g[tid * ldg + tid];
        chol_inplace(g, n, ldg, dinv, diag0, colj, 1e-13);
        tri_inverse_upper(g, n, ldg, dinv);
        apply_inverse(xs, m8, n4, n8, lds, g, ldg, dinv);
    }
    const int nt = n8 >> 3, ntiles = nt * (nt + 1) / 2;
    for (int p = 0; p < passes; ++p) {
        // Gram (lower triangle) in float64 on the tensor cores
        for (int id = warp; id < ntiles; id += kOrthThreads >> 5) {
            int ti = 0, rem = id;
            while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
            const int tj = ti + rem;
            const int ca = 8 * ti + fr, cb = 8 * tj + fr;
            const bool va = ca < lds, vb = cb < lds;
            const float* pa = xs + (size_t)fc * lds + (va ? ca : 0);
            const float* pb = xs + (size_t)fc * lds + (vb ? cb : 0);
            double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
            for (int r0 = 0; r0 < m8; r0 += 8) {
                const double a0 = va ? (double)pa[(size_t)r0 * lds] : 0.0, b0 = vb ? (double)pb[(size_t)r0 * lds] : 0.0;
                const double a1 = va ? (double)pa[(size_t)(r0 + 4) * lds] : 0.0, b1 = vb ? (double)pb[(size_t)(r0 + 4) * lds] : 0.0;
                orth_dmma(c0, c1, a0, b0);
                orth_dmma(d0, d1, a1, b1);
            }
            c0 += d0;
            c1 += d1;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 8 * ti + fr, j = 8 * tj + 2 * fc + e;
                const double v = e ? c1 : c0;
                if (j >= i) g[j * ldg + i] = v;
                else g[i * ldg + j] = v;
            }
        }
        __syncthreads();
        if (tid < n) diag0[tid] = g[tid * ldg + tid];
        // first pass: relative to the column's own squared norm; later passes: columns are unit or exactly zero
        chol_inplace(g, n, ldg, dinv, diag0, colj, p == 0 ? 1e-13 : 1e-6);
        tri_inverse_upper(g, n, ldg, dinv);
        apply_inverse(xs, m8, n4, n8, lds, g, ldg, dinv);
    }
    for (int idx = tid; idx < m * nq; idx += kOrthThreads) {
        const int r = idx / nq, q = idx - r * nq;
        const float4 v = *reinterpret_cast<const float4*>(xs + (size_t)r * lds + 4 * q);
        float* dst = xb + (size_t)r * ldx + 4 * q;
        if (vec && 4 * q + 3 < n) {
            *reinterpret_cast<float4*>(dst) = v;
        } else {
            if (4 * q < n) dst[0] = v.x;
            if (4 * q + 1 < n) dst[1] = v.y;
            if (4 * q + 2 < n) dst[2] = v.z;
            if (4 * q + 3 < n) dst[3] = v.w;
        }
    }
}

}  // namespace pmd

static size_t block_orth_smem(int64_t m, int64_t n) {
    const int64_t n4 = (n + 3) & ~3ll, n8 = (n + 7) & ~7ll, m8 = (m + 7) & ~7ll;
    const int64_t ldg = (n8 & 8) ? n8 : n8 + 8;
    return (size_t)((n8 * ldg + 3 * n8) * sizeof(double) + m8 * n4 * sizeof(float));
}

extern "C" int pmd_block_orth(float* x, int64_t batch, int64_t m, int64_t n, int64_t ldx, const double* g_ext,
                              int64_t passes, void* stream) {
    const char* fn = "pmd_block_orth";
    PMD_REQUIRE(x, fn, "null pointer");
    PMD_REQUIRE(batch > 0 && m > 0 && n > 0 && n <= 64 && ldx >= n && passes >= 0 && passes <= 3, fn, "bad size (n <= 64)");
    const size_t smem = block_orth_smem(m, n);
    PMD_REQUIRE(smem <= 227 * 1024, fn, "matrix does not fit shared memory");
    cudaError_t e = cudaFuncSetAttribute(pmd::block_orth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    pmd::block_orth_kernel<<<(unsigned)batch, pmd::kOrthThreads, smem, (cudaStream_t)stream>>>(x, (int)m, (int)n, (int)ldx, g_ext,
                                                                                              (int)passes);
    return pmd::check_launch(fn);
}
