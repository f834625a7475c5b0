// K7: the full-movie temporal projection  Z = U^T ((Y - mu) / sigma)   (pmd_loader.py:316-346, 392-414).
//
// The movie is streamed once per kernel in its native frame-major layout.  U is block structured:
// every local column lives on one bh x bw tile, the background columns are dense.
//
//  project_local : one warp owns (block, group of <= 4 components, pixel slab) and keeps its slice of
//                  U (pre-divided by sigma) in REGISTERS: lane l owns block pixels l, l+32, ...  Frames are
//                  consumed 8 at a time; the 8 x 4 per-lane partial sums are reduced across the warp with
//                  a 31-shuffle transpose-reduction, so the reduction costs ~2 instructions per output.
//  project_dense : background columns (k <= 16): lane owns 4 consecutive pixels, CTA = 1024 pixels,
//                  partial sums reduced in the warp, across warps in shared memory, across pixel tiles
//                  with float atomics.
#include "common.cuh"

namespace pmd {

// transpose-reduce 32 per-lane values: afterwards lane l holds sum over lanes of v[l] (in v[0])
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
        const int h = n / 2;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < h) {
                const float send = upper ? v[k] : v[k + h];
                const float keep = upper ? v[k + h] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
    }
    return v[0];
}

constexpr int kLocWarps = 8;
constexpr int kLocFramesPerWarp = 64;

template <typename T, int PPL>
__global__ void __launch_bounds__(kLocWarps * 32)
project_local_kernel(const T* __restrict__ movie, int64_t t, int64_t d2, int64_t d, const int32_t* __restrict__ starts,
                     int bh, int bw, const int32_t* __restrict__ ranks, const int64_t* __restrict__ col0,
                     const int32_t* __restrict__ tasks, int nslab, const float* __restrict__ uvals,
                     const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ z, int64_t ldz) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int task = blockIdx.y / nslab, slab = blockIdx.y % nslab;
    const int b = tasks[2 * task], cstart = tasks[2 * task + 1];
    const int rk = ranks[b];
    const int nc = min(4, rk - cstart);
    const int bpix = bh * bw;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int64_t colbase = col0[b] + cstart;

    float uw[4][PPL];
    float mu[PPL];
    int off[PPL];
#pragma unroll
    for (int i = 0; i < PPL; ++i) {
        const int q = slab * 32 * PPL + lane + 32 * i;
        const bool ok = q < bpix;
        int o = 0;
        float m = 0.f, is = 0.f;
        if (ok) {
            const int qi = q / bw, qj = q - qi * bw;
            o = (i0 + qi) * (int)d2 + j0 + qj;
            m = mean ? mean[o] : 0.f;
            is = inv_std ? inv_std[o] : 1.f;
        }
        off[i] = o;
        mu[i] = m;
#pragma unroll
        for (int c = 0; c < 4; ++c) uw[c][i] = (ok && c < nc) ? uvals[(colbase + c) * bpix + q] * is : 0.f;
    }

    const int64_t fbase = ((int64_t)blockIdx.x * kLocWarps + warp) * kLocFramesPerWarp;
    for (int64_t fg = fbase; fg < min(fbase + kLocFramesPerWarp, t); fg += 8) {
        float acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int64_t f = fg + g;
            if (f < t) {
                const T* fr = movie + f * d;
#pragma unroll
                for (int i = 0; i < PPL; ++i) {
                    const float x = to_f32(fr[off[i]]) - mu[i];
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[g * 4 + c] = fmaf(uw[c][i], x, acc[g * 4 + c]);
                }
            }
        }
        const float tot = transpose_reduce32(acc, lane);
        const int g = lane >> 2, c = lane & 3;
        const int64_t f = fg + g;
        if (c < nc && f < t) {
            float* dst = z + (colbase + c) * ldz + f;
            if (nslab == 1) *dst = tot;
            else atomicAdd(dst, tot);
        }
    }
}

constexpr int kDenseWarps = 8;
constexpr int kDenseFrames = 256;  // frames per CTA
constexpr int kDenseBatch = 8;     // 2-frame steps buffered between block reductions

template <typename T>
__global__ void __launch_bounds__(kDenseWarps * 32)
project_dense_kernel(const T* __restrict__ movie, int64_t t, int64_t d, const float* __restrict__ basis, int k,
                     const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ z, int64_t ldz) {
    __shared__ float red[kDenseWarps][kDenseBatch][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p0 = ((int64_t)blockIdx.x * kDenseWarps + warp) * 128 + lane * 4;
    float bs[16][4];
    float mu[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ok[j] = (p0 + j) < d;
        mu[j] = (ok[j] && mean) ? mean[p0 + j] : 0.f;
        const float is = ok[j] ? (inv_std ? inv_std[p0 + j] : 1.f) : 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) bs[c][j] = (ok[j] && c < k) ? basis[(int64_t)c * d + p0 + j] * is : 0.f;
    }
    const int64_t f_begin = (int64_t)blockIdx.y * kDenseFrames;
    const int64_t f_end = min(f_begin + kDenseFrames, t);
    for (int64_t fb = f_begin; fb < f_end; fb += 2 * kDenseBatch) {
#pragma unroll
        for (int sp = 0; sp < kDenseBatch / 2; ++sp) {
            // issue the loads of 4 frames up front (memory-level parallelism), then two 2-frame reductions
            float xs[4][4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int64_t f = fb + 4 * sp + g;
                const T* fr = movie + f * d + p0;
#pragma unroll
                for (int j = 0; j < 4; ++j) xs[g][j] = (ok[j] && f < f_end) ? to_f32(fr[j]) - mu[j] : 0.f;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float acc[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) acc[q] = 0.f;
#pragma unroll
                for (int g = 0; g < 2; ++g)
#pragma unroll
                    for (int c = 0; c < 16; ++c)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[g * 16 + c] = fmaf(bs[c][j], xs[2 * h + g][j], acc[g * 16 + c]);
                red[warp][2 * sp + h][lane] = transpose_reduce32(acc, lane);
            }
        }
        __syncthreads();
        {
            const int st = threadIdx.x >> 5, q = threadIdx.x & 31;  // 8 steps x 32 outputs = 256 threads
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < kDenseWarps; ++w) tot += red[w][st][q];
            const int g = q >> 4, c = q & 15;
            const int64_t f = fb + 2 * st + g;
            if (c < k && f < f_end) atomicAdd(&z[(int64_t)c * ldz + f], tot);
        }
        __syncthreads();
    }
}

}  // namespace pmd

extern "C" int pmd_project_local(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* starts,
                                 int64_t nb, int64_t bh, int64_t bw, const int32_t* ranks, const int64_t* col0,
                                 const int32_t* tasks, int64_t n_tasks, const float* uvals32, const float* mean,
                                 const float* inv_std, float* z, int64_t ldz, void* stream) {
    const char* fn = "pmd_project_local";
    PMD_REQUIRE(movie && starts && ranks && col0 && tasks && uvals32 && z, fn, "null pointer");
    PMD_REQUIRE(t > 0 && nb > 0 && n_tasks > 0 && ldz >= t, fn, "bad size");
    PMD_REQUIRE(d < (int64_t)1 << 31, fn, "frame too large for 32-bit pixel offsets");
    const int64_t bpix = bh * bw;
    // pixels per lane: smallest register tile that covers the block with the fewest slabs
    int ppl;
    if (bpix <= 32 * 4) ppl = 4;
    else if (bpix <= 32 * 8) ppl = 8;
    else if (bpix <= 32 * 13) ppl = 13;
    else if (bpix <= 32 * 16) ppl = 16;
    else ppl = (bpix % (32 * 13) == 0 || (bpix + 32 * 13 - 1) / (32 * 13) <= (bpix + 32 * 16 - 1) / (32 * 16)) ? 13 : 16;
    const int64_t nslab = (bpix + 32 * ppl - 1) / (32 * ppl);
    PMD_REQUIRE(n_tasks * nslab <= 65535 * 16, fn, "too many tasks");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t fper = (int64_t)pmd::kLocWarps * pmd::kLocFramesPerWarp;
    // gridDim.y is limited to 65535: launch in slices of tasks
    const int64_t max_tasks_per_launch = 65535 / nslab;
    for (int64_t t0 = 0; t0 < n_tasks; t0 += max_tasks_per_launch) {
        const int64_t nt = std::min(max_tasks_per_launch, n_tasks - t0);
        dim3 grid((unsigned)((t + fper - 1) / fper), (unsigned)(nt * nslab));
#define PMD_LAUNCH_LOCAL(PPL)                                                                                         \
    pmd::project_local_kernel<scalar_t, PPL><<<grid, pmd::kLocWarps * 32, 0, st>>>(                                  \
        (const scalar_t*)movie, t, d2, d, starts, (int)bh, (int)bw, ranks, col0, tasks + 2 * t0, (int)nslab, uvals32, \
        mean, inv_std, z, ldz)
        PMD_DISPATCH_DTYPE(dtype, fn, {
            switch (ppl) {
                case 4: PMD_LAUNCH_LOCAL(4); break;
                case 8: PMD_LAUNCH_LOCAL(8); break;
                case 13: PMD_LAUNCH_LOCAL(13); break;
                default: PMD_LAUNCH_LOCAL(16); break;
            }
        });
#undef PMD_LAUNCH_LOCAL
        int rc = pmd::check_launch(fn);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int pmd_project_dense(const void* movie, int dtype, int64_t t, int64_t d, const float* basis, int64_t k,
                                 const float* mean, const float* inv_std, float* z, int64_t ldz, void* stream) {
    const char* fn = "pmd_project_dense";
    PMD_REQUIRE(movie && basis && z, fn, "null pointer");
    PMD_REQUIRE(t > 0 && d > 0 && k > 0 && k <= 16 && ldz >= t, fn, "bad size (k <= 16)");
    const int64_t ptiles = (d + pmd::kDenseWarps * 128 - 1) / (pmd::kDenseWarps * 128);
    const int64_t fsplits = (t + pmd::kDenseFrames - 1) / pmd::kDenseFrames;
    PMD_REQUIRE(fsplits <= 65535, fn, "too many frames per call");
    dim3 grid((unsigned)ptiles, (unsigned)fsplits);
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        pmd::project_dense_kernel<scalar_t><<<grid, pmd::kDenseWarps * 32, 0, st>>>((const scalar_t*)movie, t, d, basis, (int)k,
                                                                                    mean, inv_std, z, ldz);
    });
    return pmd::check_launch(fn);
}

// =================================================================================================
// K7 v2: supertile projection.
//
// v1 (project_local_kernel above) lets every (block, component group) warp fetch its own 20x20 window
// from global memory, so every movie element crosses L2 -> SM about 4 x 1.6 times.  Here a CTA owns a
// "supertile" of G x G neighbouring blocks and walks a frame range: per sub-tile of `ft` frames the
// union of the blocks' windows (e.g. 40 x 40 pixels for G = 3, 20 x 20 blocks) is staged ONCE into
// shared memory -- already centred and scaled, (y - mu) * (1/sigma) -- and then every warp applies the
// U slice it keeps in registers (lane l owns block pixels l, l+32, ...) to all staged frames, 8 frames
// per transpose-reduction.  Re-read factor from L2 drops from ~6.4 to (S / (S - overlap))^2 ~ 1.8, all
// global loads are full row segments of the region, and the standardisation is done once per element.
// Tasks = (block, group of <= 4 components); a warp normally owns exactly one task for the whole frame
// range, so its U slice is loaded once per CTA.
// =================================================================================================
namespace pmd {

constexpr int kSTWarps = 16;
constexpr int kSTThreads = kSTWarps * 32;
constexpr int kSTPixPerThread = 4;  // staged region <= 2048 pixels

template <typename T, int PPL>
__global__ void __launch_bounds__(kSTThreads, 1)
project_supertile_kernel(const T* __restrict__ movie, int64_t t, int64_t d2, int64_t d, const int4* __restrict__ tiles,
                         const int32_t* __restrict__ task_ptr, const int4* __restrict__ tasks, int bh, int bw,
                         const float* __restrict__ uvals, const float* __restrict__ mean, const float* __restrict__ inv_std,
                         float* __restrict__ z, int64_t ldz, int ft, int frames_per_cta, int rwp, int rp) {
    extern __shared__ __align__(16) float stile[];  // [ft][rp]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int4 tl = tiles[blockIdx.x];
    const int r0 = tl.x, c0 = tl.y, rh = tl.z, rw = tl.w;
    const int npx = rh * rw;
    const int bpix = bh * bw;

    float smu[kSTPixPerThread], sis[kSTPixPerThread];
    int goff[kSTPixPerThread], soff[kSTPixPerThread];
    bool pv[kSTPixPerThread];
#pragma unroll
    for (int j = 0; j < kSTPixPerThread; ++j) {
        const int p = tid + kSTThreads * j;
        pv[j] = p < npx;
        const int r = pv[j] ? p / rw : 0;
        const int c = pv[j] ? p - r * rw : 0;
        goff[j] = (r0 + r) * (int)d2 + c0 + c;
        soff[j] = r * rwp + c;
        smu[j] = (pv[j] && mean) ? mean[goff[j]] : 0.f;
        sis[j] = (pv[j] && inv_std) ? inv_std[goff[j]] : 1.f;
    }

    const int tbeg = task_ptr[blockIdx.x], tend = task_ptr[blockIdx.x + 1];
    int cached = -1, col = 0, nc = 0;
    float uw[4][PPL];
    int off[PPL];

    const int64_t fbase = (int64_t)blockIdx.y * frames_per_cta;
    const int64_t fend = min(t, fbase + frames_per_cta);
    for (int64_t fs = fbase; fs < fend; fs += ft) {
        const int nf = (int)min((int64_t)ft, fend - fs);
        // ---- stage: centred + scaled region, one row segment per warp-load
#pragma unroll 4
        for (int f = 0; f < nf; ++f) {
            const T* fr = movie + (fs + f) * d;
#pragma unroll
            for (int j = 0; j < kSTPixPerThread; ++j)
                if (pv[j]) stile[f * rp + soff[j]] = (to_f32(fr[goff[j]]) - smu[j]) * sis[j];
        }
        __syncthreads();
        // ---- compute
        for (int ti = tbeg + warp; ti < tend; ti += kSTWarps) {
            if (ti != cached) {
                const int4 tk = tasks[ti];
                col = tk.z;
                nc = tk.w;
#pragma unroll
                for (int i = 0; i < PPL; ++i) {
                    const int q = lane + 32 * i;
                    const bool ok = q < bpix;
                    const int qi = ok ? q / bw : 0;
                    const int qj = ok ? q - qi * bw : 0;
                    off[i] = (tk.x + qi) * rwp + tk.y + qj;
#pragma unroll
                    for (int c = 0; c < 4; ++c) uw[c][i] = (ok && c < nc) ? uvals[(int64_t)(col + c) * bpix + q] : 0.f;
                }
                cached = ti;
            }
            for (int fg = 0; fg < nf; fg += 8) {
                float acc[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[k] = 0.f;
                if (nc > 2) {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (fg + g < nf) {
                            const float* fr = stile + (fg + g) * rp;
#pragma unroll
                            for (int i = 0; i < PPL; ++i) {
                                const float x = fr[off[i]];
#pragma unroll
                                for (int c = 0; c < 4; ++c) acc[g * 4 + c] = fmaf(uw[c][i], x, acc[g * 4 + c]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (fg + g < nf) {
                            const float* fr = stile + (fg + g) * rp;
#pragma unroll
                            for (int i = 0; i < PPL; ++i) {
                                const float x = fr[off[i]];
                                acc[g * 4 + 0] = fmaf(uw[0][i], x, acc[g * 4 + 0]);
                                acc[g * 4 + 1] = fmaf(uw[1][i], x, acc[g * 4 + 1]);
                            }
                        }
                    }
                }
                const float tot = transpose_reduce32(acc, lane);
                const int g = lane >> 2, c = lane & 3;
                if (c < nc && fg + g < nf) z[(int64_t)(col + c) * ldz + fs + fg + g] = tot;
            }
        }
        __syncthreads();
    }
}

}  // namespace pmd

extern "C" int pmd_project_supertile(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* tiles,
                                     int64_t n_tiles, const int32_t* task_ptr, const int32_t* tasks, int64_t bh, int64_t bw,
                                     int64_t max_region_h, int64_t max_region_w, const float* uvals32, const float* mean,
                                     const float* inv_std, float* z, int64_t ldz, void* stream) {
    const char* fn = "pmd_project_supertile";
    PMD_REQUIRE(movie && tiles && task_ptr && tasks && uvals32 && z, fn, "null pointer");
    PMD_REQUIRE(t > 0 && n_tiles > 0 && ldz >= t, fn, "bad size");
    PMD_REQUIRE(d < (int64_t)1 << 31, fn, "frame too large for 32-bit pixel offsets");
    PMD_REQUIRE(max_region_h * max_region_w <= pmd::kSTThreads * pmd::kSTPixPerThread, fn, "region larger than 2048 pixels");
    const int64_t bpix = bh * bw;
    PMD_REQUIRE(bpix <= 32 * 16, fn, "block larger than 512 pixels (use pmd_project_local)");
    const int ppl = bpix <= 32 * 4 ? 4 : bpix <= 32 * 8 ? 8 : bpix <= 32 * 13 ? 13 : 16;
    const int rwp = (int)max_region_w + 1;
    const int rp = (int)max_region_h * rwp;
    int ft = (int)((220 * 1024) / ((size_t)rp * sizeof(float)));
    ft = std::min(64, ft / 8 * 8);
    PMD_REQUIRE(ft >= 8, fn, "region does not fit shared memory");
    const size_t smem = (size_t)ft * rp * sizeof(float);
    const int frames_per_cta = std::max(ft, 512 / ft * ft);
    const int64_t fsplits = (t + frames_per_cta - 1) / frames_per_cta;
    PMD_REQUIRE(fsplits <= 65535, fn, "too many frames per call");
    dim3 grid((unsigned)n_tiles, (unsigned)fsplits);
    cudaStream_t st = (cudaStream_t)stream;
#define PMD_LAUNCH_ST(PPL)                                                                                               \
    {                                                                                                                    \
        auto k = pmd::project_supertile_kernel<scalar_t, PPL>;                                                           \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }         \
        k<<<grid, pmd::kSTThreads, smem, st>>>((const scalar_t*)movie, t, d2, d, (const int4*)tiles, task_ptr,            \
                                               (const int4*)tasks, (int)bh, (int)bw, uvals32, mean, inv_std, z, ldz, ft, \
                                               frames_per_cta, rwp, rp);                                                 \
    }
    PMD_DISPATCH_DTYPE(dtype, fn, {
        switch (ppl) {
            case 4: PMD_LAUNCH_ST(4); break;
            case 8: PMD_LAUNCH_ST(8); break;
            case 13: PMD_LAUNCH_ST(13); break;
            default: PMD_LAUNCH_ST(16); break;
        }
    });
#undef PMD_LAUNCH_ST
    return pmd::check_launch(fn);
}
