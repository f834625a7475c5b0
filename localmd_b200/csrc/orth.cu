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

namespace pmd {

constexpr int kOrthThreads = 256;

// in-place Cholesky of the n x n float64 matrix g (row-major, ld), lower triangle; dead[j] = 1 marks dropped
// columns (their row/column of L is zero and 1/L_jj is taken as 0).  tol: relative pivot threshold.
__device__ void chol_inplace(double* g, int n, int ld, double* dinv, const double* diag0, double tol) {
    const int tid = threadIdx.x;
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        const double piv = g[j * ld + j];
        const bool ok = piv > tol * diag0[j] && piv > 0.0;
        const double ljj = ok ? sqrt(piv) : 0.0;
        const double inv = ok ? 1.0 / ljj : 0.0;
        __syncthreads();
        if (tid == 0) {
            g[j * ld + j] = ljj;
            dinv[j] = inv;
        }
        for (int i = j + 1 + tid; i < n; i += kOrthThreads) g[i * ld + j] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: g[i][k] -= l[i][j] * l[k][j], j < k <= i < n
        const int rem = n - j - 1;
        for (int idx = tid; idx < rem * rem; idx += kOrthThreads) {
            const int a = idx / rem, b = idx - a * rem;
            if (b <= a) g[(j + 1 + a) * ld + j + 1 + b] -= g[(j + 1 + a) * ld + j] * g[(j + 1 + b) * ld + j];
        }
    }
    __syncthreads();
}

// X <- X L^-T (row-wise forward substitution), X in shared memory [m][lds] float32, L lower (float64)
__device__ void solve_rows(float* xs, int m, int n, int lds, const double* l, int ld, const double* dinv) {
    for (int r = threadIdx.x; r < m; r += kOrthThreads) {
        float* x = xs + (size_t)r * lds;
        for (int j = 0; j < n; ++j) {
            double acc = (double)x[j];
            for (int k = 0; k < j; ++k) acc -= (double)x[k] * l[j * ld + k];
            x[j] = (float)(acc * dinv[j]);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kOrthThreads)
block_orth_kernel(float* __restrict__ x, int m, int n, int ldx, const double* __restrict__ g_ext, int passes) {
    extern __shared__ __align__(16) unsigned char osm[];
    const int ldg = n | 1;
    double* g = reinterpret_cast<double*>(osm);          // [n][ldg]
    double* dinv = g + (size_t)n * ldg;                  // [n]
    double* diag0 = dinv + n;                            // [n]
    const int lds = n | 1;                               // odd row stride: conflict-free column walks
    float* xs = reinterpret_cast<float*>(diag0 + n);     // [m][lds]
    const int tid = threadIdx.x;
    float* xb = x + (size_t)blockIdx.x * m * ldx;

    for (int idx = tid; idx < m * n; idx += kOrthThreads) {
        const int r = idx / n, c = idx - r * n;
        xs[(size_t)r * lds + c] = xb[(size_t)r * ldx + c];
    }
    if (g_ext) {
        const double* ge = g_ext + (size_t)blockIdx.x * n * n;
        for (int idx = tid; idx < n * n; idx += kOrthThreads) g[(idx / n) * ldg + idx % n] = ge[idx];
        __syncthreads();
        if (tid < n) diag0[tid] = g[tid * ldg + tid];
        chol_inplace(g, n, ldg, dinv, diag0, 1e-13);
        solve_rows(xs, m, n, lds, g, ldg, dinv);
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p) {
        // Gram (lower triangle) in float64
        const int npair = n * (n + 1) / 2;
        for (int idx = tid; idx < npair; idx += kOrthThreads) {
            int i = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
            while ((i + 1) * (i + 2) / 2 <= idx) ++i;
            while (i * (i + 1) / 2 > idx) --i;
            const int j = idx - i * (i + 1) / 2;          // j <= i
            double acc = 0.0;
            for (int r = 0; r < m; ++r) acc = fma((double)xs[(size_t)r * lds + i], (double)xs[(size_t)r * lds + j], acc);
            g[i * ldg + j] = acc;
        }
        __syncthreads();
        if (tid < n) diag0[tid] = g[tid * ldg + tid];
        // first pass: relative to the column's own squared norm; later passes: columns are unit or exactly zero
        chol_inplace(g, n, ldg, dinv, diag0, p == 0 ? 1e-13 : 1e-6);
        solve_rows(xs, m, n, lds, g, ldg, dinv);
    }
    for (int idx = tid; idx < m * n; idx += kOrthThreads) {
        const int r = idx / n, c = idx - r * n;
        xb[(size_t)r * ldx + c] = xs[(size_t)r * lds + c];
    }
}

}  // namespace pmd

extern "C" int pmd_block_orth(float* x, int64_t batch, int64_t m, int64_t n, int64_t ldx, const double* g_ext,
                              int64_t passes, void* stream) {
    const char* fn = "pmd_block_orth";
    PMD_REQUIRE(x, fn, "null pointer");
    PMD_REQUIRE(batch > 0 && m > 0 && n > 0 && n <= 64 && ldx >= n && passes >= 0 && passes <= 3, fn, "bad size (n <= 64)");
    const size_t smem = ((size_t)n * (n | 1) + 2 * n) * sizeof(double) + (size_t)m * (n | 1) * sizeof(float);
    PMD_REQUIRE(smem <= 227 * 1024, fn, "matrix does not fit shared memory");
    cudaError_t e = cudaFuncSetAttribute(pmd::block_orth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    pmd::block_orth_kernel<<<(unsigned)batch, pmd::kOrthThreads, smem, (cudaStream_t)stream>>>(x, (int)m, (int)n, (int)ldx, g_ext,
                                                                                              (int)passes);
    return pmd::check_launch(fn);
}
