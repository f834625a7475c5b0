// K7 v3: full-movie projection  z = U^T ((Y - mean) / std)  as ONE streaming pass (pmd_loader.py:316-346,
// 392-414: v_projection / v_projection_routine).
//
// A CTA owns a COLUMN STRIP of the field of view (G neighbouring block columns, all rows) and a tile of
// 256 frames, and walks the strip top to bottom one pixel row at a time:
//   * staging: the row segment (<= 48 pixels) of all 256 frames is loaded with coalesced loads,
//     centred/scaled once, and written TRANSPOSED into shared memory as x[pixel][frame] (16-byte frame
//     chunks XOR-swizzled by the pixel index, so both the transposing STS.128 and the LDS.128 of the
//     compute phase are bank-conflict free).  Two buffers: the loads of row i+1 are in flight while row i
//     is consumed.
//   * compute: every warp owns one TASK = (block, group of <= 8 of its components) for the 20 (bh) rows
//     the block spans -- or 8 of the dense background components restricted to the strip's own columns,
//     for all rows -- with an 8 components x 8 frames register tile per lane (lane l owns frames 4l..4l+3
//     and 128+4l..).  The task's U values of the current row sit in a small per-warp slab as
//     [pixel][8 comps] and are read as broadcast LDS.128.  When a block ends the warp stores its
//     z rows and picks up the next task of its slot (the host packs tasks into slots, ops.make_strips).
// Every movie element is read from L2/HBM (1 + 1/G) times (horizontal halo only), every U value once
// per 256 frames, and nothing is reduced across lanes or with atomics: a local z element is written
// by exactly one lane; background columns produce one partial row set per strip (summed by the host).
#include "common.cuh"

namespace pmd {

constexpr int kPSWarps = 8;
constexpr int kPSThreads = kPSWarps * 32;
constexpr int kPSF = 256;          // frames per CTA
constexpr int kPSMaxRW = 48;       // widest strip (pixels)
constexpr int kPSPhases = 4;       // a row is staged / consumed in this many interleaved phases
constexpr int kPSUnits = 2;        // staging units (pixel x 4 frames) per thread, pixel set and phase

struct PSItem {                    // one CTA column: 8 ints
    int c0, rw, task_ptr, n_rows, bg_part, row0, pad1, pad2;
};
struct PSTask {                    // 12 ints
    int by, bx, h, w, col, nc, ncp, urow, uoff_lo, uoff_hi, kind, pad;
};

__device__ __forceinline__ void ps_cp_async16(void* smem, const void* gmem) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(gmem));
}

// SETS = ceil(strip width / 32): lane l stages pixels l (+32); warp w stages the 4-frame chunks w, w+8, ...
// FULL: every frame of the tile exists (the host launches the last, partial tile separately with FULL = false).
template <typename T, int SETS, bool FULL>
__global__ void __launch_bounds__(kPSThreads, 2)
project_stream_kernel(const T* __restrict__ movie, int64_t t, int64_t tile0, int64_t d2, int64_t d, const PSItem* __restrict__ items,
                      const int32_t* __restrict__ slot_ptr, const PSTask* __restrict__ tasks, const float* __restrict__ upack,
                      const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ z, int64_t ldz,
                      float* __restrict__ zbg, int64_t ldzbg, int64_t bg_stride) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PSItem it = items[blockIdx.x];
    const int rw = it.rw, c0 = it.c0, row0 = it.row0, row_end = it.row0 + it.n_rows;
    // the last tile is shifted back so that it is full (it recomputes, and rewrites with identical values, a few
    // frames of its neighbour); only movies shorter than one tile take the clamped FULL = false path
    const int64_t f0 = FULL ? min((tile0 + blockIdx.y) * kPSF, t - kPSF) : (tile0 + blockIdx.y) * kPSF;
    const int xstride = rw * kPSF;
    float* const ubuf = sm + 2 * xstride + warp * (2 * kPSMaxRW * 8);   // two U slabs per warp

    // ---- staging -------------------------------------------------------------------------------------
    // unit (s, m): pixel k = lane + 32 s, frames f0 + 4 c4 .. +3 with c4 = warp + 8 m (m = 0..7); phase ph handles
    // m = 2 ph, 2 ph + 1.  Frames past the end of the movie are CLAMPED to the last frame: their results are never
    // stored, so no predicate or zero fill is needed on the data path.
    T pre[SETS][kPSUnits][4];
    const T* const col_ptr = movie + c0 + lane;            // + row * d2 + frame * d
    const int64_t fbase = f0 + 4 * warp;                   // first frame of unit m = 0
    const int64_t ustep = 32 * d;                          // frames advance by 32 from unit to unit
    int dst_off[SETS];
#pragma unroll
    for (int s = 0; s < SETS; ++s) {
        const int k = lane + 32 * s;
        dst_off[s] = k * kPSF + ((warp ^ (k & 7)) << 2);   // + 32 m  (c4 ^ (k & 7) = (warp ^ (k & 7)) + 8 m)
    }
    auto prefetch = [&](const T* rowp, int ph) {            // raw loads only: nothing here waits for them
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
            if (lane + 32 * s < rw) {
#pragma unroll
                for (int n = 0; n < kPSUnits; ++n) {
                    const int m = kPSUnits * ph + n;
                    if (FULL) {
                        const T* src = rowp + 32 * s + fbase * d + (int64_t)m * ustep;
#pragma unroll
                        for (int j = 0; j < 4; ++j) pre[s][n][j] = src[(int64_t)j * d];
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int64_t f = min(fbase + 32 * m + j, t - 1);
                            pre[s][n][j] = rowp[32 * s + f * d];
                        }
                    }
                }
            }
        }
    };
    float mu[SETS], is[SETS];
    auto load_norm = [&](int row) {
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
            mu[s] = 0.f;
            is[s] = 1.f;
            if (lane + 32 * s < rw) {
                const int64_t pix = (int64_t)row * d2 + c0 + lane + 32 * s;
                if (mean) mu[s] = __ldg(mean + pix);
                if (inv_std) is[s] = __ldg(inv_std + pix);
            }
        }
    };
    auto commit = [&](float* dst, int ph) {                 // centre, scale, transpose into x[pixel][frame]
#pragma unroll
        for (int s = 0; s < SETS; ++s) {
            if (lane + 32 * s < rw) {
#pragma unroll
                for (int n = 0; n < kPSUnits; ++n) {
                    const int m = kPSUnits * ph + n;
                    float4 v;
                    v.x = (to_f32(pre[s][n][0]) - mu[s]) * is[s];
                    v.y = (to_f32(pre[s][n][1]) - mu[s]) * is[s];
                    v.z = (to_f32(pre[s][n][2]) - mu[s]) * is[s];
                    v.w = (to_f32(pre[s][n][3]) - mu[s]) * is[s];
                    *reinterpret_cast<float4*>(dst + dst_off[s] + 32 * m) = v;
                }
            }
        }
    };

    // ---- task state of this warp ---------------------------------------------------------------------
    const int sp = it.task_ptr + warp;   // slot_ptr entries of this item: [task_ptr .. task_ptr + kPSWarps]
    int tnext = slot_ptr[sp];
    const int tend = slot_ptr[sp + 1];
    PSTask tk;
    tk.by = 1 << 30;
    tk.h = 0;
    if (tnext < tend) tk = tasks[tnext];
    // U slab of task row r: [w][ncp] floats, contiguous in upack; fetched one row ahead with cp.async
    auto fetch_slab = [&](int uoff_lo, int uoff_hi, int urow, int n4, int r, int buf) {
        const int64_t uo = ((int64_t)uoff_hi << 32 | (uint32_t)uoff_lo) + (int64_t)r * urow;
        const float4* usrc = reinterpret_cast<const float4*>(upack + uo);
        float4* udst = reinterpret_cast<float4*>(ubuf + buf * (kPSMaxRW * 8));
        for (int i = lane; i < n4; i += 32) ps_cp_async16(udst + i, usrc + i);
    };
    int ub = 0;                                            // slab buffer holding the current row of the current task
    if (tk.by == row0) fetch_slab(tk.uoff_lo, tk.uoff_hi, tk.urow, (tk.w * tk.ncp) >> 2, 0, 0);
    asm volatile("cp.async.commit_group;\n" ::);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const T* rowp = col_ptr + (int64_t)row0 * d2;
    load_norm(row0);
#pragma unroll
    for (int ph = 0; ph < kPSPhases; ++ph) {
        prefetch(rowp, ph);
        commit(sm, ph);
    }
    __syncthreads();

    for (int row = row0; row < row_end; ++row) {
        const float* xs = sm + ((row - row0) & 1) * xstride;
        float* xn = sm + ((row - row0 + 1) & 1) * xstride;
        const bool more = row + 1 < row_end;
        rowp += d2;
        if (more) load_norm(row + 1);
        const bool active = row >= tk.by && row < tk.by + tk.h;
        const int w = tk.w;
        const float* us = ubuf + ub * (kPSMaxRW * 8);
        // the slab of this row was requested one row ago (or before the loop); request the next one
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncwarp();
        if (active && row + 1 < tk.by + tk.h) {
            fetch_slab(tk.uoff_lo, tk.uoff_hi, tk.urow, (w * tk.ncp) >> 2, row + 1 - tk.by, ub ^ 1);
        } else if (!active && row + 1 == tk.by) {
            fetch_slab(tk.uoff_lo, tk.uoff_hi, tk.urow, (w * tk.ncp) >> 2, 0, ub);   // starts next row: buffer is free
        }
        asm volatile("cp.async.commit_group;\n" ::);
#pragma unroll
        for (int ph = 0; ph < kPSPhases; ++ph) {
            if (more) prefetch(rowp, ph);
            if (active) {
                const int j0 = (w * ph) / kPSPhases, j1 = (w * (ph + 1)) / kPSPhases;
                const float* xr = xs + (tk.bx + j0) * kPSF;
                if (tk.ncp == 8) {
#pragma unroll 1
                    for (int j = j0; j < j1; ++j, xr += kPSF) {
                        const int sw = (lane ^ ((tk.bx + j) & 7)) << 2;
                        const float4 xa = *reinterpret_cast<const float4*>(xr + sw);
                        const float4 xb = *reinterpret_cast<const float4*>(xr + 128 + sw);
                        const float4 u0 = *reinterpret_cast<const float4*>(us + j * 8);
                        const float4 u1 = *reinterpret_cast<const float4*>(us + j * 8 + 4);
                        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                        const float uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                        for (int c = 0; c < 8; ++c)
#pragma unroll
                            for (int f = 0; f < 8; ++f) acc[c][f] = fmaf(uv[c], xv[f], acc[c][f]);
                    }
                } else {
#pragma unroll 1
                    for (int j = j0; j < j1; ++j, xr += kPSF) {
                        const int sw = (lane ^ ((tk.bx + j) & 7)) << 2;
                        const float4 xa = *reinterpret_cast<const float4*>(xr + sw);
                        const float4 xb = *reinterpret_cast<const float4*>(xr + 128 + sw);
                        const float4 u0 = *reinterpret_cast<const float4*>(us + j * 4);
                        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                        const float uv[4] = {u0.x, u0.y, u0.z, u0.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int f = 0; f < 8; ++f) acc[c][f] = fmaf(uv[c], xv[f], acc[c][f]);
                    }
                }
            }
            if (more) commit(xn, ph);
        }
        if (active) ub ^= 1;
        if (active && row == tk.by + tk.h - 1) {
            // the block (or the strip, for a background task) ends here: store and take the next task
            float* zo;
            int64_t ldo;
            if (tk.kind == 0) {
                zo = z + (int64_t)tk.col * ldz;
                ldo = ldz;
            } else {
                zo = zbg + (int64_t)it.bg_part * bg_stride + (int64_t)tk.col * ldzbg;
                ldo = ldzbg;
            }
            const bool vec = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(zo) & 15) == 0);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < tk.nc) {
                    float* o = zo + (int64_t)c * ldo;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int64_t f = f0 + 128 * hh + 4 * lane;
                        if (vec && f + 3 < t) {
                            *reinterpret_cast<float4*>(o + f) =
                                make_float4(acc[c][4 * hh], acc[c][4 * hh + 1], acc[c][4 * hh + 2], acc[c][4 * hh + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (f + j < t) o[f + j] = acc[c][4 * hh + j];
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            ++tnext;
            if (tnext < tend) {
                tk = tasks[tnext];
                if (tk.by == row + 1) {                      // back-to-back tasks: the slab request above was skipped
                    fetch_slab(tk.uoff_lo, tk.uoff_hi, tk.urow, (tk.w * tk.ncp) >> 2, 0, ub);
                    asm volatile("cp.async.commit_group;\n" ::);
                }
            } else {
                tk.by = 1 << 30;
                tk.h = 0;
            }
        }
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
}

}  // namespace pmd

extern "C" int pmd_project_stream(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* items,
                                  int64_t n_items, const int32_t* slot_ptr, const int32_t* tasks, int64_t max_rw,
                                  const float* upack, const float* mean, const float* inv_std, float* z, int64_t ldz,
                                  float* zbg, int64_t ldzbg, int64_t bg_stride, void* stream) {
    const char* fn = "pmd_project_stream";
    PMD_REQUIRE(movie && items && slot_ptr && tasks && upack && z && zbg, fn, "null pointer");
    PMD_REQUIRE(t > 0 && n_items > 0 && ldz >= t && ldzbg >= t, fn, "bad size");
    PMD_REQUIRE(max_rw > 0 && max_rw <= pmd::kPSMaxRW, fn, "strip wider than 48 pixels");
    PMD_REQUIRE(((uintptr_t)upack & 15) == 0, fn, "upack must be 16-byte aligned");
    const int64_t ftiles = (t + pmd::kPSF - 1) / pmd::kPSF;
    PMD_REQUIRE(ftiles <= 65535, fn, "too many frames per call");
    const size_t smem = (size_t)(2 * max_rw * pmd::kPSF + pmd::kPSWarps * 2 * pmd::kPSMaxRW * 8) * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t full_tiles = t >= pmd::kPSF ? ftiles : 0;   // t >= 256: every tile is full (the last one overlaps)
#define PMD_LAUNCH_PS(SETS, FULL, TILE0, NT)                                                                             \
    if ((NT) > 0) {                                                                                                      \
        auto k = pmd::project_stream_kernel<scalar_t, SETS, FULL>;                                                       \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }         \
        k<<<dim3((unsigned)n_items, (unsigned)(NT)), pmd::kPSThreads, smem, st>>>(                                       \
            (const scalar_t*)movie, t, (int64_t)(TILE0), d2, d, (const pmd::PSItem*)items, slot_ptr,                     \
            (const pmd::PSTask*)tasks, upack, mean, inv_std, z, ldz, zbg, ldzbg, bg_stride);                             \
    }
    PMD_DISPATCH_DTYPE(dtype, fn, {
        if (max_rw <= 32) {
            PMD_LAUNCH_PS(1, true, 0, full_tiles)
            PMD_LAUNCH_PS(1, false, full_tiles, ftiles - full_tiles)
        } else {
            PMD_LAUNCH_PS(2, true, 0, full_tiles)
            PMD_LAUNCH_PS(2, false, full_tiles, ftiles - full_tiles)
        }
    });
#undef PMD_LAUNCH_PS
    return pmd::check_launch(fn);
}
