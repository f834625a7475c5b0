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

}  // namespace pmd

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
