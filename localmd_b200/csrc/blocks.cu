// Per-block kernels of the local decomposition stage (decomposition.py:192-330, 811-853 and
// evaluation.py:84-222 of the reference), batched over ALL blocks of the field of view at once.
// The reference runs a serial, host-synchronised Python loop over blocks (decomposition.py:790-838);
// here every step is one launch over (block, tile) and the host only reads back ranks[nb].
#include "common.cuh"

namespace pmd {

// ------------------------------------------------------------------------------------------------
// pooling geometry (XLA 'SAME' padding with stride == window == saf)
// ------------------------------------------------------------------------------------------------
struct PoolGeom {
    int ph, pw, lo_h, lo_w;
};
__host__ __device__ inline PoolGeom pool_geom(int bh, int bw, int saf) {
    PoolGeom g;
    g.ph = (bh + saf - 1) / saf;
    g.pw = (bw + saf - 1) / saf;
    g.lo_h = (g.ph * saf - bh) / 2;
    g.lo_w = (g.pw * saf - bw) / 2;
    return g;
}

__global__ void block_unpool_kernel(const float* __restrict__ uds, int64_t nb, int bh, int bw, int saf, int r, int rp,
                                    float* __restrict__ w) {
    const PoolGeom g = pool_geom(bh, bw, saf);
    const int P = g.ph * g.pw;
    const int bpix = bh * bw;
    const int64_t total = nb * (int64_t)bpix * rp;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % rp);
        const int64_t bq = idx / rp;
        const int q = (int)(bq % bpix);
        const int64_t b = bq / bpix;
        float v = 0.f;
        if (c < r) {
            const int qi = q / bw, qj = q % bw;
            const int pi = (qi + g.lo_h) / saf, pj = (qj + g.lo_w) / saf;
            const int r0 = max(pi * saf - g.lo_h, 0), r1 = min(pi * saf - g.lo_h + saf, bh);
            const int c0 = max(pj * saf - g.lo_w, 0), c1 = min(pj * saf - g.lo_w + saf, bw);
            v = uds[(b * P + pi * g.pw + pj) * r + c] / (float)((r1 - r0) * (c1 - c0));
        }
        w[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// roughness statistics + rank rule
// ------------------------------------------------------------------------------------------------
__global__ void spatial_stat_kernel(const float* __restrict__ u, int bh, int bw, int r, int rp, float* __restrict__ sstat) {
    const int64_t b = blockIdx.x;
    const int bpix = bh * bw;
    const float* ub = u + b * (int64_t)bpix * rp;
    for (int c = threadIdx.x; c < r; c += blockDim.x) {
        double sd = 0.0, sa = 0.0;
        for (int qi = 0; qi < bh; ++qi) {
            for (int qj = 0; qj < bw; ++qj) {
                const float x = ub[(qi * bw + qj) * rp + c];
                sa += (double)fabsf(x);
                if (qi + 1 < bh) sd += (double)fabsf(ub[((qi + 1) * bw + qj) * rp + c] - x);
                if (qj + 1 < bw) sd += (double)fabsf(x - ub[(qi * bw + qj + 1) * rp + c]);
            }
        }
        const float cnt = (float)((bh - 1) * bw + bh * (bw - 1));
        const float avg_diff = (float)sd / cnt;
        const float avg_elem = (float)(sa / (double)bpix);
        sstat[b * r + c] = avg_diff / avg_elem;
    }
}

__global__ void __launch_bounds__(128) temporal_stat_kernel(const float* __restrict__ v, int64_t nrows, int64_t t, int64_t ldv,
                                                            float* __restrict__ tstat) {
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const float* vr = v + row * ldv;
    double sd = 0.0, sa = 0.0;
    for (int64_t i = lane; i < t; i += 32) {
        const float m = vr[i];
        sa += (double)fabsf(m);
        if (i + 2 < t) sd += (double)fabsf((m + vr[i + 2]) - 2.f * vr[i + 1]);
    }
    sd = warp_sum(sd);
    sa = warp_sum(sa);
    if (lane == 0) {
        const float num = (float)(sd / (double)(t - 2));
        const float den = (float)(sa / (double)t);
        tstat[row] = num / den;
    }
}

__global__ void rank_select_kernel(const float* __restrict__ sstat, const float* __restrict__ tstat, int64_t nb, int r,
                                   float thr_s, float thr_t, int max_fail, int32_t* __restrict__ ranks) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    int fails = 0, kept = 0;
    for (int c = 0; c < r; ++c) {
        const bool good = (sstat[b * r + c] < thr_s) && (tstat[b * r + c] < thr_t);  // NaN compares false
        ++kept;  // a failing component is still kept until the failure budget is exhausted
        if (!good) {
            if (++fails == max_fail) break;
        } else {
            fails = 0;
        }
    }
    ranks[b] = kept;
}

__global__ void assemble_u_kernel(const float* __restrict__ u, int bh, int bw, int rp, const int32_t* __restrict__ starts,
                                  const int32_t* __restrict__ ranks, const int64_t* __restrict__ col0,
                                  const float* __restrict__ block_weights, const double* __restrict__ cumw, int64_t d2,
                                  double* __restrict__ uvals64, float* __restrict__ uvals32) {
    const int64_t b = blockIdx.x;
    const int bpix = bh * bw;
    const int rk = ranks[b];
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    for (int idx = threadIdx.x; idx < rk * bpix; idx += blockDim.x) {
        const int c = idx / bpix, q = idx % bpix;
        const int qi = q / bw, qj = q % bw;
        const int64_t pix = (int64_t)(i0 + qi) * d2 + j0 + qj;
        const double val = (1.0 / cumw[pix]) * ((double)u[(b * bpix + q) * rp + c] * (double)block_weights[q]);
        const int64_t o = (col0[b] + c) * bpix + q;
        uvals64[o] = val;
        uvals32[o] = (float)val;
    }
}

}  // namespace pmd

// =================================================================================================
extern "C" int pmd_block_unpool(const float* uds, int64_t nb, int64_t bh, int64_t bw, int64_t saf, int64_t r, int64_t rp,
                                float* w, void* stream) {
    const char* fn = "pmd_block_unpool";
    PMD_REQUIRE(uds && w, fn, "null pointer");
    PMD_REQUIRE(nb > 0 && r > 0 && rp >= r && rp % 4 == 0, fn, "bad size (rp multiple of 4, >= r)");
    const int64_t total = nb * bh * bw * rp;
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    pmd::block_unpool_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(uds, nb, (int)bh, (int)bw, (int)saf, (int)r, (int)rp, w);
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_stats_rank(const float* u, const float* v, int64_t nb, int64_t bh, int64_t bw, int64_t r,
                                    int64_t rp, int64_t t, int64_t ldv, float thr_s, float thr_t, int64_t max_fail,
                                    float* sstat, float* tstat, int32_t* ranks, void* stream) {
    const char* fn = "pmd_block_stats_rank";
    PMD_REQUIRE(u && v && sstat && tstat && ranks, fn, "null pointer");
    PMD_REQUIRE(nb > 0 && r > 0 && rp >= r && t > 2 && ldv >= t && max_fail >= 1 && bh > 1 && bw > 1, fn, "bad size");
    cudaStream_t st = (cudaStream_t)stream;
    pmd::spatial_stat_kernel<<<(unsigned)nb, 64, 0, st>>>(u, (int)bh, (int)bw, (int)r, (int)rp, sstat);
    const int64_t nrows = nb * r;
    pmd::temporal_stat_kernel<<<(unsigned)((nrows + 3) / 4), 128, 0, st>>>(v, nrows, t, ldv, tstat);
    pmd::rank_select_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(sstat, tstat, nb, (int)r, thr_s, thr_t, (int)max_fail,
                                                                          ranks);
    return pmd::check_launch(fn);
}

extern "C" int pmd_assemble_u(const float* u, int64_t nb, int64_t bh, int64_t bw, int64_t rp, const int32_t* starts,
                              const int32_t* ranks, const int64_t* col0, const float* block_weights, const double* cumw,
                              int64_t d2, double* uvals64, float* uvals32, void* stream) {
    const char* fn = "pmd_assemble_u";
    PMD_REQUIRE(u && starts && ranks && col0 && block_weights && cumw && uvals64 && uvals32, fn, "null pointer");
    PMD_REQUIRE(nb > 0, fn, "bad size");
    pmd::assemble_u_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(u, (int)bh, (int)bw, (int)rp, starts, ranks, col0,
                                                                          block_weights, cumw, d2, uvals64, uvals32);
    return pmd::check_launch(fn);
}
