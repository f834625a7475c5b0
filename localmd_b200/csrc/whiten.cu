// Sparse Gram of the local columns of U for the whitening step  G = M^T (U^T U) M
// (decomposition.py:974-996).  The reference forms U^T U with scipy.sparse on the host; here the
// block structure is used directly: two local columns interact only when their block windows
// overlap, so U^T U restricted to the local columns is a block-sparse matrix with one dense
// rank(b1) x rank(b2) tile per ordered pair of overlapping blocks.  One CTA per pair, float64 (FP64 tensor cores).
#include <vector>

#include "common.cuh"

namespace pmd {

// Tile (b1, b2) = U_b1^T U_b2 over the overlap rectangle of the two block windows, on the FP64 tensor cores: per image row of
// the overlap and 4 pixels, a warp loads one A fragment (8 components of b1 x 4 pixels) and one B fragment per 8 components
// of b2 -- every value is read once per warp row instead of once per output entry (the scalar version spent 25 ms on the
// 23 k pairs of the C4 shard, 17 x 17 entries x ~700 pixels each).
constexpr int kUtpMaxNT = 13;   // ceil(102 / 8) column fragments

__global__ void __launch_bounds__(128)
utu_pairs_kernel(const int32_t* __restrict__ pairs, const int64_t* __restrict__ pair_rowoff, const int32_t* __restrict__ starts,
                 int bh, int bw, const int32_t* __restrict__ ranks, const int64_t* __restrict__ col0,
                 const double* __restrict__ uvals, const int64_t* __restrict__ rowptr, double* __restrict__ vals,
                 int32_t* __restrict__ cols) {
    const int64_t p = blockIdx.x;
    const int b1 = pairs[2 * p], b2 = pairs[2 * p + 1];
    const int r1 = ranks[b1], r2 = ranks[b2];
    if (r1 == 0 || r2 == 0) return;
    const int i1 = starts[2 * b1], j1 = starts[2 * b1 + 1], i2 = starts[2 * b2], j2 = starts[2 * b2 + 1];
    const int ilo = max(i1, i2), ihi = min(i1, i2) + bh, jlo = max(j1, j2), jhi = min(j1, j2) + bw;
    const int bpix = bh * bw;
    const int64_t c01 = col0[b1], c02 = col0[b2];
    const int64_t ro = pair_rowoff[p];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int nt2 = (r2 + 7) >> 3;
    for (int ti = warp; 8 * ti < r1; ti += 4) {
        const int c1 = 8 * ti + fr;
        const bool v1 = c1 < r1;
        const double* u1 = uvals + (c01 + (v1 ? c1 : 0)) * bpix - (int64_t)i1 * bw - j1;
        double acc[kUtpMaxNT][2];
#pragma unroll
        for (int n = 0; n < kUtpMaxNT; ++n) acc[n][0] = acc[n][1] = 0.0;
        for (int i = ilo; i < ihi; ++i) {
            for (int jb = jlo; jb < jhi; jb += 4) {
                const int j = jb + fc;
                const bool vj = j < jhi;
                const double a = (v1 && vj) ? __ldg(u1 + (int64_t)i * bw + j) : 0.0;
#pragma unroll
                for (int n = 0; n < kUtpMaxNT; ++n) {
                    if (n < nt2) {
                        const int c2 = 8 * n + fr;
                        const double b = (vj && c2 < r2) ? __ldg(uvals + (c02 + c2) * bpix + (int64_t)(i - i2) * bw + (j - j2)) : 0.0;
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                                     : "+d"(acc[n][0]), "+d"(acc[n][1])
                                     : "d"(a), "d"(b));
                    }
                }
            }
        }
        if (v1) {
            const int64_t base = rowptr[c01 + c1] + ro;
#pragma unroll
            for (int n = 0; n < kUtpMaxNT; ++n) {
                if (n < nt2) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c2 = 8 * n + 2 * fc + e;
                        if (c2 < r2) {
                            vals[base + c2] = acc[n][e];
                            cols[base + c2] = (int32_t)(c02 + c2);
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Z_loc = (U_loc^T U_loc) X for a dense right factor X [n_local][m] (float64), from the dense rank(b1) x rank(b2) tiles
// that utu_pairs_kernel wrote -- on the FP64 tensor cores (mma.sync m8n8k4).  A CTA owns the rows of one block b1 and 256
// columns of X (a warp: 32 columns = 4 fragments, 24 rows = 3 fragments at a time) and walks the (<= 9) blocks b2 that
// overlap b1: the rows of X that belong to b2 are read ONCE per pair as B fragments, the tile is read as A fragments
// through L1.  The CSR product this replaces gathered a 13 KB row of X per stored entry: 6 GB of L2 traffic at C2
// (2.1 ms) and 88 GB at the C4 shard (mean rank 17: 55 ms, the largest item of that job).
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kUtaThreads = 256;

__global__ void __launch_bounds__(kUtaThreads)
utu_apply_tiles_kernel(const int32_t* __restrict__ pairs, const int32_t* __restrict__ seg_ptr, const int64_t* __restrict__ pair_rowoff,
                       const int32_t* __restrict__ ranks, const int64_t* __restrict__ col0, const int64_t* __restrict__ rowptr,
                       const double* __restrict__ vals, const double* __restrict__ x, int64_t ldx, int m, double* __restrict__ z,
                       int64_t ldz) {
    const int b1 = blockIdx.y;
    const int r1 = ranks[b1];
    if (r1 == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int j0 = blockIdx.x * kUtaThreads + warp * 32;
    if (j0 >= m) return;
    const int64_t c01 = col0[b1];
    const int p0 = seg_ptr[b1], p1 = seg_ptr[b1 + 1];
    for (int row0 = 0; row0 < r1; row0 += 24) {
        double acc[3][4][2];
        int64_t rp[3];
        bool rv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            rv[i] = row0 + 8 * i + fr < r1;
            rp[i] = rv[i] ? rowptr[c01 + row0 + 8 * i + fr] : 0;
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;
        }
        for (int p = p0; p < p1; ++p) {
            const int b2 = pairs[2 * p + 1];
            const int r2 = ranks[b2];
            const int64_t ro = pair_rowoff[p];
            const double* xb = x + col0[b2] * ldx;
            for (int k0 = 0; k0 < r2; k0 += 4) {
                const int k = k0 + fc;
                const bool kv = k < r2;
                double a[3], bfr[4];
#pragma unroll
                for (int i = 0; i < 3; ++i) a[i] = (rv[i] && kv) ? __ldg(vals + rp[i] + ro + k) : 0.0;
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const int j = j0 + 8 * n + fr;
                    bfr[n] = (kv && j < m) ? __ldg(xb + (int64_t)k * ldx + j) : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int n = 0; n < 4; ++n)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                                     : "+d"(acc[i][n][0]), "+d"(acc[i][n][1])
                                     : "d"(a[i]), "d"(bfr[n]));
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (!rv[i]) continue;
            double* zr = z + (c01 + row0 + 8 * i + fr) * ldz;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int j = j0 + 8 * n + 2 * fc;
                if (j < m) zr[j] = acc[i][n][0];
                if (j + 1 < m) zr[j + 1] = acc[i][n][1];
            }
        }
    }
}

}  // namespace pmd

extern "C" int pmd_utu_pairs(const int32_t* pairs, int64_t n_pairs, const int64_t* pair_rowoff, const int32_t* starts,
                             int64_t bh, int64_t bw, const int32_t* ranks, const int64_t* col0, const double* uvals64,
                             const int64_t* rowptr, double* vals, int32_t* cols, void* stream) {
    const char* fn = "pmd_utu_pairs";
    PMD_REQUIRE(pairs && pair_rowoff && starts && ranks && col0 && uvals64 && rowptr && vals && cols, fn, "null pointer");
    PMD_REQUIRE(n_pairs > 0 && n_pairs < ((int64_t)1 << 31) && bh > 0 && bw > 0, fn, "bad size");
    pmd::utu_pairs_kernel<<<(unsigned)n_pairs, 128, 0, (cudaStream_t)stream>>>(pairs, pair_rowoff, starts, (int)bh, (int)bw, ranks,
                                                                              col0, uvals64, rowptr, vals, cols);
    return pmd::check_launch(fn);
}


// Host bookkeeping of pmd_utu_pairs (plain C++, no device work; the Python driver calls it on its worker thread, where a
// ctypes call runs without the interpreter lock): for the block pairs sorted by (b1, b2) and the kept ranks,
//   pair_rowoff[p] = number of entries that precede tile p in each of its rows,  rowptr = CSR row pointer of U_loc^T U_loc.
extern "C" int pmd_utu_host_tables(const int32_t* pairs, int64_t n_pairs, const int64_t* ranks, int64_t nb, int64_t* pair_rowoff,
                                   int64_t* rowptr) {
    const char* fn = "pmd_utu_host_tables";
    PMD_REQUIRE(pairs && ranks && pair_rowoff && rowptr, fn, "null pointer");
    PMD_REQUIRE(n_pairs >= 0 && nb > 0, fn, "bad size");
    std::vector<int64_t> width((size_t)nb, 0);
    int64_t run = 0, prev = -1;
    for (int64_t p = 0; p < n_pairs; ++p) {
        const int64_t b1 = pairs[2 * p], b2 = pairs[2 * p + 1];
        PMD_REQUIRE(b1 >= 0 && b1 < nb && b2 >= 0 && b2 < nb && b1 >= prev, fn, "pairs must be sorted by b1 and index blocks");
        if (b1 != prev) { run = 0; prev = b1; }
        pair_rowoff[p] = run;
        run += ranks[b2];
        width[(size_t)b1] += ranks[b2];
    }
    int64_t row = 0, acc = 0;
    rowptr[0] = 0;
    for (int64_t b = 0; b < nb; ++b)
        for (int64_t c = 0; c < ranks[b]; ++c) { acc += width[(size_t)b]; rowptr[++row] = acc; }
    return 0;
}

extern "C" int pmd_utu_apply_tiles(const int32_t* pairs, const int32_t* seg_ptr, int64_t nb, const int64_t* pair_rowoff,
                                   const int32_t* ranks, const int64_t* col0, const int64_t* rowptr, const double* vals,
                                   const double* x, int64_t ldx, int64_t m, double* z, int64_t ldz, void* stream) {
    const char* fn = "pmd_utu_apply_tiles";
    PMD_REQUIRE(pairs && seg_ptr && pair_rowoff && ranks && col0 && rowptr && vals && x && z, fn, "null pointer");
    PMD_REQUIRE(nb > 0 && nb <= 65535 && m > 0 && m < ((int64_t)1 << 31) && ldx >= m && ldz >= m, fn, "bad size (nb <= 65535)");
    dim3 grid((unsigned)((m + pmd::kUtaThreads - 1) / pmd::kUtaThreads), (unsigned)nb);
    pmd::utu_apply_tiles_kernel<<<grid, pmd::kUtaThreads, 0, (cudaStream_t)stream>>>(pairs, seg_ptr, pair_rowoff, ranks, col0, rowptr, vals, x,
                                                                                    ldx, (int)m, z, ldz);
    return pmd::check_launch(fn);
}
