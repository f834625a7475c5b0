// CSR export of the sparse spatial matrix U (decomposition.py:811-857 coo -> csr, 912-933 background columns):
// the canonical CSR (rows = pixels, ascending columns, exact zeros dropped) is written DIRECTLY from the block-component
// form -- a pixel's entries are the kept components of the (few) blocks that cover it, in block order (= column order),
// followed by the dense background columns -- instead of materialising 8.6 M coordinate triplets and sorting them
// (the torch path: ~45 launches, a 64-bit radix sort and three host synchronisations, 3.4 ms of device time at C2).
// Two passes of one thread per pixel: count (then an inclusive scan by the caller), fill.
#include "common.cuh"

namespace pmd {

template <bool FILL>
__global__ void __launch_bounds__(256)
export_csr_kernel(const double* __restrict__ uvals, const float* __restrict__ bg, int K, int d1, int d2,
                  const int32_t* __restrict__ row_starts, int n_br, const int32_t* __restrict__ col_starts, int n_bc, int bh, int bw,
                  const int32_t* __restrict__ ranks, const int64_t* __restrict__ col0, int64_t n_local,
                  const int64_t* __restrict__ row_ids, int64_t* __restrict__ counts_rel, int64_t* __restrict__ counts_phys,
                  const int64_t* __restrict__ indptr_rel, const int64_t* __restrict__ indptr_phys, int32_t* __restrict__ cols_rel,
                  double* __restrict__ vals_rel, int32_t* __restrict__ cols_phys, float* __restrict__ vals_phys) {
    const int64_t d = (int64_t)d1 * d2;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d) return;
    const int i = (int)(p / d2), j = (int)(p - (int64_t)i * d2);
    const int64_t rid = row_ids ? row_ids[p] : p;
    int64_t n = 0, o_rel = 0, o_phys = 0;
    if (FILL) {
        o_rel = indptr_rel[rid];
        o_phys = indptr_phys[p];
    }
    const int bpix = bh * bw;
    // block rows / columns whose range contains (i, j); the start lists are ascending
    int r_lo = 0, c_lo = 0;
    while (r_lo < n_br && row_starts[r_lo] + bh <= i) ++r_lo;
    while (c_lo < n_bc && col_starts[c_lo] + bw <= j) ++c_lo;
    for (int ri = r_lo; ri < n_br && row_starts[ri] <= i; ++ri) {
        for (int ci = c_lo; ci < n_bc && col_starts[ci] <= j; ++ci) {
            const int b = ri * n_bc + ci;
            const int q = (i - row_starts[ri]) * bw + (j - col_starts[ci]);
            const int rk = ranks[b];
            const int64_t c0 = col0[b];
            for (int c = 0; c < rk; ++c) {
                const double v = uvals[(c0 + c) * bpix + q];
                if (v != 0.0) {
                    if (FILL) {
                        cols_rel[o_rel + n] = (int32_t)(c0 + c);
                        vals_rel[o_rel + n] = v;
                        cols_phys[o_phys + n] = (int32_t)(c0 + c);
                        vals_phys[o_phys + n] = (float)v;
                    }
                    ++n;
                }
            }
        }
    }
    for (int k = 0; k < K; ++k) {
        const float v = bg[(int64_t)k * d + p];
        if (v != 0.f) {
            if (FILL) {
                cols_rel[o_rel + n] = (int32_t)(n_local + k);
                vals_rel[o_rel + n] = (double)v;
                cols_phys[o_phys + n] = (int32_t)(n_local + k);
                vals_phys[o_phys + n] = v;
            }
            ++n;
        }
    }
    if (!FILL) {
        counts_rel[rid] = n;
        counts_phys[p] = n;
    }
}

}  // namespace pmd

extern "C" int pmd_export_csr(const double* uvals, const float* bg, int64_t K, int64_t d1, int64_t d2, const int32_t* row_starts,
                              int64_t n_br, const int32_t* col_starts, int64_t n_bc, int64_t bh, int64_t bw, const int32_t* ranks,
                              const int64_t* col0, int64_t n_local, const int64_t* row_ids, int fill, int64_t* counts_rel,
                              int64_t* counts_phys, const int64_t* indptr_rel, const int64_t* indptr_phys, int32_t* cols_rel,
                              double* vals_rel, int32_t* cols_phys, float* vals_phys, void* stream) {
    const char* fn = "pmd_export_csr";
    PMD_REQUIRE(d1 > 0 && d2 > 0 && d1 * d2 < (1ll << 31) && bh > 0 && bw > 0 && K >= 0 && n_br >= 0 && n_bc >= 0, fn, "bad size");
    PMD_REQUIRE(n_local + K < (1ll << 31), fn, "column ids must fit int32");
    PMD_REQUIRE((K == 0 || bg) && (n_br * n_bc == 0 || (uvals && row_starts && col_starts && ranks && col0)), fn, "null pointer");
    const int64_t d = d1 * d2;
    const unsigned grid = (unsigned)((d + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (!fill) {
        PMD_REQUIRE(counts_rel && counts_phys, fn, "null pointer");
        pmd::export_csr_kernel<false><<<grid, 256, 0, st>>>(uvals, bg, (int)K, (int)d1, (int)d2, row_starts, (int)n_br, col_starts, (int)n_bc,
                                                            (int)bh, (int)bw, ranks, col0, n_local, row_ids, counts_rel, counts_phys,
                                                            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    } else {
        PMD_REQUIRE(indptr_rel && indptr_phys && cols_rel && vals_rel && cols_phys && vals_phys, fn, "null pointer");
        pmd::export_csr_kernel<true><<<grid, 256, 0, st>>>(uvals, bg, (int)K, (int)d1, (int)d2, row_starts, (int)n_br, col_starts, (int)n_bc,
                                                           (int)bh, (int)bw, ranks, col0, n_local, row_ids, nullptr, nullptr, indptr_rel,
                                                           indptr_phys, cols_rel, vals_rel, cols_phys, vals_phys);
    }
    return pmd::check_launch(fn);
}
