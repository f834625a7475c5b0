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
constexpr int kPSPhases = 3;       // a row is staged / consumed in this many interleaved phases
constexpr int kPSUnits = 4;        // staging units (pixel x 4 frames) per thread and phase

struct PSItem {                    // one CTA column: 8 ints
    int c0, rw, task_ptr, n_rows, bg_part, row0, pad1, pad2;
};
struct PSTask {                    // 12 ints
    int by, bx, h, w, col, nc, ncp, urow, uoff_lo, uoff_hi, kind, pad;
};

template <typename T>
__global__ void __launch_bounds__(kPSThreads, 2)
project_stream_kernel(const T* __restrict__ movie, int64_t t, int64_t d2, int64_t d, const PSItem* __restrict__ items,
                      const int32_t* __restrict__ slot_ptr, const PSTask* __restrict__ tasks, const float* __restrict__ upack,
                      const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ z, int64_t ldz,
                      float* __restrict__ zbg, int64_t ldzbg, int64_t bg_stride) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const PSItem it = items[blockIdx.x];
    const int rw = it.rw, c0 = it.c0, row_end = it.row0 + it.n_rows;
    const int64_t f0 = (int64_t)blockIdx.y * kPSF;
    float* xbuf[2] = {sm, sm + rw * kPSF};
    float* ubuf = sm + 2 * rw * kPSF + warp * (kPSMaxRW * 8);

    // ---- staging helpers --------------------------------------------------------------------------
    const int nunits = rw * (kPSF / 4);
    T pre[kPSUnits][4];
    auto prefetch = [&](int row, int ph) {   // raw loads only: nothing here waits for them
#pragma unroll
        for (int n = 0; n < kPSUnits; ++n) {
            const int idx = tid + kPSThreads * (ph * kPSUnits + n);
            if (idx < nunits) {
                const int c4 = idx / rw, k = idx - c4 * rw;
                const T* src = movie + (f0 + 4 * c4) * d + (int64_t)row * d2 + c0 + k;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t f = f0 + 4 * c4 + j;
                    pre[n][j] = f < t ? src[(int64_t)j * d] : T(0);
                }
            }
        }
    };
    auto commit = [&](float* dst, int row, int ph) {   // centre, scale, transpose into x[pixel][frame]
#pragma unroll
        for (int n = 0; n < kPSUnits; ++n) {
            const int idx = tid + kPSThreads * (ph * kPSUnits + n);
            if (idx < nunits) {
                const int c4 = idx / rw, k = idx - c4 * rw;
                const int64_t pix = (int64_t)row * d2 + c0 + k;
                const float mu = mean ? __ldg(mean + pix) : 0.f;
                const float is = inv_std ? __ldg(inv_std + pix) : 1.f;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (f0 + 4 * c4 + j < t) ? (to_f32(pre[n][j]) - mu) * is : 0.f;
                *reinterpret_cast<float4*>(dst + k * kPSF + ((c4 ^ (k & 7)) << 2)) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    };

    // ---- task state of this warp --------------------------------------------------------------------
    const int sp = it.task_ptr + warp;   // slot_ptr entries of this item: [task_ptr .. task_ptr + kPSWarps]
    int tnext = slot_ptr[sp], tend = slot_ptr[sp + 1];
    PSTask tk;
    tk.by = 1 << 30;
    tk.h = 0;
    if (tnext < tend) tk = tasks[tnext];
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int ph = 0; ph < kPSPhases; ++ph) {
        prefetch(it.row0, ph);
        commit(xbuf[0], it.row0, ph);
    }
    __syncthreads();

    for (int row = it.row0; row < row_end; ++row) {
        const float* xs = xbuf[(row - it.row0) & 1];
        float* xn = xbuf[(row - it.row0 + 1) & 1];
        const bool more = row + 1 < row_end;
        const bool active = row >= tk.by && row < tk.by + tk.h;
        const int w = tk.w, bx = tk.bx;
        if (active) {
            // U slab of this row: [w][ncp] floats, contiguous in upack
            const int64_t uo = ((int64_t)tk.uoff_hi << 32 | (uint32_t)tk.uoff_lo) + (int64_t)(row - tk.by) * tk.urow;
            const float4* usrc = reinterpret_cast<const float4*>(upack + uo);
            const int n4 = (w * tk.ncp) >> 2;
            for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(ubuf)[i] = usrc[i];
            __syncwarp();
        }
#pragma unroll
        for (int ph = 0; ph < kPSPhases; ++ph) {
            if (more) prefetch(row + 1, ph);
            if (active) {
                const int j0 = (w * ph) / kPSPhases, j1 = (w * (ph + 1)) / kPSPhases;
                if (tk.ncp == 8) {
#pragma unroll 1
                    for (int j = j0; j < j1; ++j) {
                        const int k = bx + j;
                        const float* xr = xs + k * kPSF;
                        const int sw = (lane ^ (k & 7)) << 2;
                        const float4 xa = *reinterpret_cast<const float4*>(xr + sw);
                        const float4 xb = *reinterpret_cast<const float4*>(xr + 128 + sw);
                        const float4 u0 = *reinterpret_cast<const float4*>(ubuf + j * 8);
                        const float4 u1 = *reinterpret_cast<const float4*>(ubuf + j * 8 + 4);
                        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                        const float uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                        for (int c = 0; c < 8; ++c)
#pragma unroll
                            for (int f = 0; f < 8; ++f) acc[c][f] = fmaf(uv[c], xv[f], acc[c][f]);
                    }
                } else {
#pragma unroll 1
                    for (int j = j0; j < j1; ++j) {
                        const int k = bx + j;
                        const float* xr = xs + k * kPSF;
                        const int sw = (lane ^ (k & 7)) << 2;
                        const float4 xa = *reinterpret_cast<const float4*>(xr + sw);
                        const float4 xb = *reinterpret_cast<const float4*>(xr + 128 + sw);
                        const float4 u0 = *reinterpret_cast<const float4*>(ubuf + j * 4);
                        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                        const float uv[4] = {u0.x, u0.y, u0.z, u0.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int f = 0; f < 8; ++f) acc[c][f] = fmaf(uv[c], xv[f], acc[c][f]);
                    }
                }
            }
            if (more) commit(xn, row + 1, ph);
        }
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
            } else {
                tk.by = 1 << 30;
                tk.h = 0;
            }
        }
        __syncthreads();
    }
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
    const size_t smem = (size_t)(2 * max_rw * pmd::kPSF + pmd::kPSWarps * pmd::kPSMaxRW * 8) * sizeof(float);
    dim3 grid((unsigned)n_items, (unsigned)ftiles);
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        auto k = pmd::project_stream_kernel<scalar_t>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
        k<<<grid, pmd::kPSThreads, smem, st>>>((const scalar_t*)movie, t, d2, d, (const pmd::PSItem*)items, slot_ptr,
                                               (const pmd::PSTask*)tasks, upack, mean, inv_std, z, ldz, zbg, ldzbg, bg_stride);
    });
    return pmd::check_launch(fn);
}
