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

constexpr int kPoolTau = 8;

__global__ void __launch_bounds__(128)
block_pool_tavg_kernel(const float* __restrict__ yres, int64_t t, int64_t d2, int64_t d, const int32_t* __restrict__ starts,
                       int bh, int bw, int saf, int taf, float* __restrict__ bta) {
    const PoolGeom g = pool_geom(bh, bw, saf);
    const int P = g.ph * g.pw;
    const int64_t tp = t / taf;
    const int64_t b = blockIdx.y;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int64_t tau0 = (int64_t)blockIdx.x * kPoolTau;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        const int pi = p / g.pw, pj = p % g.pw;
        const int r0 = max(pi * saf - g.lo_h, 0), r1 = min(pi * saf - g.lo_h + saf, bh);
        const int c0 = max(pj * saf - g.lo_w, 0), c1 = min(pj * saf - g.lo_w + saf, bw);
        const float cnt = (float)((r1 - r0) * (c1 - c0));
        for (int64_t tau = tau0; tau < min(tau0 + kPoolTau, tp); ++tau) {
            float acc = 0.f;
            for (int ff = 0; ff < taf; ++ff) {
                const float* fr = yres + (tau * taf + ff) * d + (int64_t)i0 * d2 + j0;
                float sum = 0.f;
                for (int r = r0; r < r1; ++r)
                    for (int c = c0; c < c1; ++c) sum += fr[(int64_t)r * d2 + c];
                acc += sum / cnt;
            }
            bta[(b * tp + tau) * P + p] = acc / (float)taf;
        }
    }
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
// out[b][c][f] = sum_q w[b][q][c] * Y_b[q][f]       (SGEMM-style: 4 comps x 4 frames per thread,
// pixel chunks of 64 staged through shared memory; threads = 16 frame groups x rp/4 comp groups)
// ------------------------------------------------------------------------------------------------
constexpr int kProjFT = 64;   // frames per CTA
constexpr int kProjKC = 64;   // pixels per staged chunk
constexpr int kProjLd = kProjFT + 4;

__global__ void block_project_kernel(const float* __restrict__ movie, int64_t mbs, int64_t t, int64_t d2, int64_t d,
                                     const int32_t* __restrict__ starts, int bh, int bw, const float* __restrict__ w,
                                     int r, int rp, float* __restrict__ out) {
    extern __shared__ __align__(16) float psm[];
    float* wch = psm;                     // [kProjKC][rp]
    float* tile = psm + kProjKC * rp;     // [kProjKC][kProjLd]
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x;
    const int ncg = rp / 4;
    const int cg = tid % ncg, fg = tid / ncg;
    const int64_t b = blockIdx.y;
    const int64_t f0 = (int64_t)blockIdx.x * kProjFT;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int bpix = bh * bw;
    const float* mv = movie + b * mbs + (int64_t)i0 * d2 + j0;
    const float* wb = w + b * (int64_t)bpix * rp;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int q0 = 0; q0 < bpix; q0 += kProjKC) {
        const int kc = min(kProjKC, bpix - q0);
        for (int idx = tid; idx < kc * rp; idx += nthreads) wch[idx] = wb[(int64_t)q0 * rp + idx];
        for (int idx = tid; idx < kProjKC * kProjFT; idx += nthreads) {
            const int ff = idx / kProjKC, qq = idx % kProjKC;
            float v = 0.f;
            if (qq < kc && f0 + ff < t) {
                const int q = q0 + qq;
                const int qi = q / bw, qj = q - qi * bw;
                v = mv[(f0 + ff) * d + (int64_t)qi * d2 + qj];
            }
            tile[qq * kProjLd + ff] = v;
        }
        __syncthreads();
        for (int qq = 0; qq < kc; ++qq) {
            const float4 wv = *reinterpret_cast<const float4*>(&wch[qq * rp + cg * 4]);
            const float4 xv = *reinterpret_cast<const float4*>(&tile[qq * kProjLd + fg * 4]);
            const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wa[i], xa[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = cg * 4 + i;
        if (c >= r) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t f = f0 + fg * 4 + j;
            if (f < t) out[(b * r + c) * t + f] = acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// s[b][q][c] = sum_f Y_b[q][f] * vb[b][c][f]        (64 pixels x rp comps per CTA, loop over time)
// ------------------------------------------------------------------------------------------------
constexpr int kSpatPT = 64;
constexpr int kSpatFC = 32;

__global__ void block_spatial_kernel(const float* __restrict__ movie, int64_t mbs, int64_t t, int64_t d2, int64_t d,
                                     const int32_t* __restrict__ starts, int bh, int bw, const float* __restrict__ vb,
                                     int r, int rp, float* __restrict__ s) {
    extern __shared__ __align__(16) float ssm[];
    float* tileT = ssm;                          // [kSpatFC][kSpatPT]
    float* vbT = ssm + kSpatFC * kSpatPT;        // [kSpatFC][rp]
    int* pixoff = reinterpret_cast<int*>(vbT + kSpatFC * rp);  // [kSpatPT]
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x;
    const int ncg = rp / 4;
    const int cg = tid % ncg, pg = tid / ncg;
    const int64_t b = blockIdx.y;
    const int q0 = blockIdx.x * kSpatPT;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int bpix = bh * bw;
    const float* mv = movie + b * mbs;
    for (int qq = tid; qq < kSpatPT; qq += nthreads) {
        const int q = q0 + qq;
        int off = -1;
        if (q < bpix) {
            const int qi = q / bw, qj = q - qi * bw;
            off = (i0 + qi) * (int)d2 + j0 + qj;
        }
        pixoff[qq] = off;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    __syncthreads();
    for (int64_t fc = 0; fc < t; fc += kSpatFC) {
        for (int idx = tid; idx < kSpatFC * kSpatPT; idx += nthreads) {
            const int ff = idx / kSpatPT, qq = idx % kSpatPT;
            const int off = pixoff[qq];
            tileT[idx] = (off >= 0 && fc + ff < t) ? mv[(fc + ff) * d + off] : 0.f;
        }
        for (int idx = tid; idx < rp * kSpatFC; idx += nthreads) {
            const int c = idx / kSpatFC, ff = idx % kSpatFC;
            vbT[ff * rp + c] = (c < r && fc + ff < t) ? vb[(b * r + c) * t + fc + ff] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int ff = 0; ff < kSpatFC; ++ff) {
            const float4 xv = *reinterpret_cast<const float4*>(&tileT[ff * kSpatPT + pg * 4]);
            const float4 vv = *reinterpret_cast<const float4*>(&vbT[ff * rp + cg * 4]);
            const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
            const float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], va[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = q0 + pg * 4 + i;
        if (q >= bpix) continue;
        *reinterpret_cast<float4*>(&s[(b * bpix + q) * rp + cg * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
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

__global__ void __launch_bounds__(128) temporal_stat_kernel(const float* __restrict__ v, int64_t nrows, int64_t t,
                                                            float* __restrict__ tstat) {
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const float* vr = v + row * t;
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
extern "C" int pmd_block_pool_tavg(const float* yres, int64_t t, int64_t d2, int64_t d, const int32_t* starts,
                                   int64_t nb, int64_t bh, int64_t bw, int64_t saf, int64_t taf, float* bta, void* stream) {
    const char* fn = "pmd_block_pool_tavg";
    PMD_REQUIRE(yres && starts && bta, fn, "null pointer");
    PMD_REQUIRE(t > 0 && nb > 0 && nb <= 65535 && bh > 0 && bw > 0 && saf > 0 && taf > 0 && t / taf > 0, fn, "bad size");
    const int64_t tp = t / taf;
    dim3 grid((unsigned)((tp + pmd::kPoolTau - 1) / pmd::kPoolTau), (unsigned)nb);
    pmd::block_pool_tavg_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(yres, t, d2, d, starts, (int)bh, (int)bw, (int)saf,
                                                                       (int)taf, bta);
    return pmd::check_launch(fn);
}

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

extern "C" int pmd_block_project(const float* movie, int64_t movie_batch_stride, int64_t t, int64_t d2, int64_t d,
                                 const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w, int64_t r,
                                 int64_t rp, float* out, void* stream) {
    const char* fn = "pmd_block_project";
    PMD_REQUIRE(movie && starts && w && out, fn, "null pointer");
    PMD_REQUIRE(t > 0 && nb > 0 && nb <= 65535 && r > 0 && rp >= r && rp % 4 == 0 && rp <= 128, fn, "bad size");
    const int threads = 16 * (int)(rp / 4);
    const size_t smem = (size_t)(pmd::kProjKC * rp + pmd::kProjKC * pmd::kProjLd) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(pmd::block_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    dim3 grid((unsigned)((t + pmd::kProjFT - 1) / pmd::kProjFT), (unsigned)nb);
    pmd::block_project_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(movie, movie_batch_stride, t, d2, d, starts,
                                                                             (int)bh, (int)bw, w, (int)r, (int)rp, out);
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_spatial(const float* movie, int64_t movie_batch_stride, int64_t t, int64_t d2, int64_t d,
                                 const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* vb, int64_t r,
                                 int64_t rp, float* s, void* stream) {
    const char* fn = "pmd_block_spatial";
    PMD_REQUIRE(movie && starts && vb && s, fn, "null pointer");
    PMD_REQUIRE(t > 0 && nb > 0 && nb <= 65535 && r > 0 && rp >= r && rp % 4 == 0 && rp <= 128, fn, "bad size");
    const int threads = 16 * (int)(rp / 4);
    const size_t smem = (size_t)(pmd::kSpatFC * pmd::kSpatPT + pmd::kSpatFC * rp) * sizeof(float) + pmd::kSpatPT * sizeof(int);
    cudaError_t e = cudaFuncSetAttribute(pmd::block_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    dim3 grid((unsigned)((bh * bw + pmd::kSpatPT - 1) / pmd::kSpatPT), (unsigned)nb);
    pmd::block_spatial_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(movie, movie_batch_stride, t, d2, d, starts,
                                                                             (int)bh, (int)bw, vb, (int)r, (int)rp, s);
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_stats_rank(const float* u, const float* v, int64_t nb, int64_t bh, int64_t bw, int64_t r,
                                    int64_t rp, int64_t t, float thr_s, float thr_t, int64_t max_fail, float* sstat,
                                    float* tstat, int32_t* ranks, void* stream) {
    const char* fn = "pmd_block_stats_rank";
    PMD_REQUIRE(u && v && sstat && tstat && ranks, fn, "null pointer");
    PMD_REQUIRE(nb > 0 && r > 0 && rp >= r && t > 2 && max_fail >= 1 && bh > 1 && bw > 1, fn, "bad size");
    cudaStream_t st = (cudaStream_t)stream;
    pmd::spatial_stat_kernel<<<(unsigned)nb, 64, 0, st>>>(u, (int)bh, (int)bw, (int)r, (int)rp, sstat);
    const int64_t nrows = nb * r;
    pmd::temporal_stat_kernel<<<(unsigned)((nrows + 3) / 4), 128, 0, st>>>(v, nrows, t, tstat);
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
