// K9: frame reconstruction behind PMDArray.__getitem__ (pmdarray.py:132-171):
//   out[n][i] = (sum_j U[pix[i]][j] * c[j][n]) * scale[pix[i]] + shift[pix[i]]
// U as CSR over physical pixel rows.  One thread = one pixel x 4 frames; consecutive threads take
// consecutive pixels so the frame-major output rows are written coalesced.  The reference instead
// materialises the dense (R_total x T) product R*diag(s)*Vt eagerly on the CPU (pmdarray.py:50-52);
// here `c` only holds the requested frames.
#include "common.cuh"

namespace pmd {

__global__ void __launch_bounds__(256)
reconstruct_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                   const float* __restrict__ c, int64_t n, const int32_t* __restrict__ pix, int64_t npix,
                   const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out) {
    const int64_t n4 = (n + 3) / 4;
    const int64_t total = npix * n4;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % npix, fq = idx / npix;
        const int64_t f0 = fq * 4;
        const int p = pix[i];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int64_t e0 = indptr[p], e1 = indptr[p + 1];
        if (f0 + 4 <= n && (n & 3) == 0) {
            for (int64_t e = e0; e < e1; ++e) {
                const float v = values[e];
                const float4 cv = *reinterpret_cast<const float4*>(&c[(int64_t)indices[e] * n + f0]);
                acc[0] = fmaf(v, cv.x, acc[0]); acc[1] = fmaf(v, cv.y, acc[1]);
                acc[2] = fmaf(v, cv.z, acc[2]); acc[3] = fmaf(v, cv.w, acc[3]);
            }
        } else {
            for (int64_t e = e0; e < e1; ++e) {
                const float v = values[e];
                const float* cr = c + (int64_t)indices[e] * n + f0;
#pragma unroll
                for (int j = 0; j < 4; ++j) if (f0 + j < n) acc[j] = fmaf(v, cr[j], acc[j]);
            }
        }
        const float sc = scale ? scale[p] : 1.f, sh = shift ? shift[p] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (f0 + j < n) out[(f0 + j) * npix + i] = fmaf(acc[j], sc, sh);
    }
}

// float64 variant used by the whitening step (U^T U M must be formed beyond float32 accuracy, see
// DESIGN.md "whitening"): out[n][i] = sum_j U[pix[i]][j] * c[j][n], everything double.
__global__ void __launch_bounds__(256)
reconstruct_f64_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                       const double* __restrict__ values, const double* __restrict__ c, int64_t n,
                       const int32_t* __restrict__ pix, int64_t npix, double* __restrict__ out) {
    const int64_t n2 = (n + 1) / 2;
    const int64_t total = npix * n2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % npix, fq = idx / npix;
        const int64_t f0 = fq * 2;
        const int p = pix[i];
        double a0 = 0.0, a1 = 0.0;
        const bool two = f0 + 1 < n;
        for (int64_t e = indptr[p]; e < indptr[p + 1]; ++e) {
            const double v = values[e];
            const double* cr = c + (int64_t)indices[e] * n + f0;
            a0 = fma(v, cr[0], a0);
            if (two) a1 = fma(v, cr[1], a1);
        }
        out[f0 * npix + i] = a0;
        if (two) out[(f0 + 1) * npix + i] = a1;
    }
}

// Z[col][f] = sum_p U[p][col] * W[f][p]  in float64.  One warp per column, 8 frames per CTA row.
// Local columns walk their block window, dense (background) columns walk all d pixels.
__global__ void __launch_bounds__(128)
project_cols_f64_kernel(const double* __restrict__ w, int64_t m, int64_t d2, int64_t d, const int32_t* __restrict__ starts,
                        int bh, int bw, const int32_t* __restrict__ blk_of_col, const int64_t* __restrict__ col0,
                        int64_t n_local, const double* __restrict__ uvals, const double* __restrict__ bg, int64_t n_cols,
                        double* __restrict__ z) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t col = (int64_t)blockIdx.y * 4 + warp;
    if (col >= n_cols) return;
    const int64_t f0 = (int64_t)blockIdx.x * 8;
    double acc[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) acc[g] = 0.0;
    if (col < n_local) {
        const int b = blk_of_col[col];
        const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
        const int bpix = bh * bw;
        const double* uc = uvals + col * bpix;
        for (int q = lane; q < bpix; q += 32) {
            const int qi = q / bw, qj = q - qi * bw;
            const int64_t off = (int64_t)(i0 + qi) * d2 + j0 + qj;
            const double u = uc[q];
#pragma unroll
            for (int g = 0; g < 8; ++g)
                if (f0 + g < m) acc[g] = fma(u, w[(f0 + g) * d + off], acc[g]);
        }
    } else {
        const double* bc = bg + (col - n_local) * d;
        for (int64_t p = lane; p < d; p += 32) {
            const double u = bc[p];
#pragma unroll
            for (int g = 0; g < 8; ++g)
                if (f0 + g < m) acc[g] = fma(u, w[(f0 + g) * d + p], acc[g]);
        }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const double tot = warp_sum(acc[g]);
        if (lane == 0 && f0 + g < m) z[col * m + f0 + g] = tot;
    }
}

}  // namespace pmd

extern "C" int pmd_reconstruct_f64(const int64_t* indptr, const int32_t* indices, const double* values, const double* c,
                                   int64_t n, const int32_t* pix, int64_t npix, double* out, void* stream) {
    const char* fn = "pmd_reconstruct_f64";
    PMD_REQUIRE(indptr && indices && values && c && pix && out, fn, "null pointer");
    PMD_REQUIRE(n >= 0 && npix >= 0, fn, "bad size");
    if (n == 0 || npix == 0) return 0;
    const int64_t total = npix * ((n + 1) / 2);
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32);
    pmd::reconstruct_f64_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(indptr, indices, values, c, n, pix, npix, out);
    return pmd::check_launch(fn);
}

extern "C" int pmd_project_cols_f64(const double* w, int64_t m, int64_t d2, int64_t d, const int32_t* starts, int64_t bh,
                                    int64_t bw, const int32_t* blk_of_col, const int64_t* col0, int64_t n_local,
                                    const double* uvals64, const double* bg64, int64_t n_cols, double* z, void* stream) {
    const char* fn = "pmd_project_cols_f64";
    PMD_REQUIRE(w && z, fn, "null pointer");
    PMD_REQUIRE(m > 0 && d > 0 && n_cols > 0 && n_local >= 0 && n_local <= n_cols, fn, "bad size");
    PMD_REQUIRE(n_local == 0 || (starts && blk_of_col && uvals64), fn, "null local-column arrays");
    PMD_REQUIRE(n_local == n_cols || bg64, fn, "null background basis");
    PMD_REQUIRE((n_cols + 3) / 4 <= 65535, fn, "too many columns");
    dim3 grid((unsigned)((m + 7) / 8), (unsigned)((n_cols + 3) / 4));
    pmd::project_cols_f64_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(w, m, d2, d, starts, (int)bh, (int)bw, blk_of_col, col0,
                                                                         n_local, uvals64, bg64, n_cols, z);
    return pmd::check_launch(fn);
}

extern "C" int pmd_reconstruct(const int64_t* indptr, const int32_t* indices, const float* values, const float* c,
                               int64_t n, const int32_t* pix, int64_t npix, const float* scale, const float* shift,
                               float* out, void* stream) {
    const char* fn = "pmd_reconstruct";
    PMD_REQUIRE(indptr && indices && values && c && pix && out, fn, "null pointer");
    PMD_REQUIRE(n >= 0 && npix >= 0, fn, "bad size");
    if (n == 0 || npix == 0) return 0;
    const int64_t total = npix * ((n + 3) / 4);
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 32);
    pmd::reconstruct_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(indptr, indices, values, c, n, pix, npix, scale, shift, out);
    return pmd::check_launch(fn);
}
