// K7 on the 5th-generation tensor cores:  z = U^T ((Y - mean) / std)  over the whole movie in ONE streaming pass
// (pmd_loader.py:316-346, 392-414: v_projection / v_projection_routine), local and background columns together.
//
// The field of view is cut into column strips that partition every image row exactly (host tables:
// strips_ts_host.cu), so every movie element is read from HBM once.  A CTA owns one strip item (strip, row range) and
// `tiles` x 128 frames, and walks the strip one image row at a time.  Per row and 32-pixel chunk, for each frame tile:
//   D[128 frames x N slot columns] += A[128 frames x 32 pixels] * B[32 pixels x N slot columns]
//   * raw movie tiles [128 frames x 32 pixels] arrive by 2-D TMA boxes (cp.async.bulk.tensor) in a ring of shared-memory
//     stages.  The grid runs strip-fastest, so the CTAs that execute together request whole image rows of the same
//     frames: measured 7.1 TB/s for these 128-byte-wide boxes (profiles/r02_tma_read_calibration.txt).
//   * two converter groups (4 warps each, thread = frame) read their frame's 32 pixels from the raw tile, subtract the
//     mean, split into hi = x with the low 13 mantissa bits cleared (exact in TF32) and a bf16 pair (bf16(hi), bf16(x - hi)),
//     and write BOTH operand tiles straight into TENSOR MEMORY (tcgen05.st): the A operand of the MMAs comes from tensor
//     memory, so the movie never goes back through shared memory (the shared-memory bandwidth is what bounded the
//     previous version, which staged A in shared memory).  1 / std is folded into the coefficients.
//   * B: the coefficient image of the strip row, prebuilt once per decomposition by pmd_pack_strips_ts in exactly the
//     shared-memory image (K-major SWIZZLE_128B; TF32 hi part + bf16 pair part (bf16(lo), bf16(hi))); one thread fetches
//     each chunk with a single bulk asynchronous copy.
//   * one thread issues, per 8 pixels, tcgen05.mma.kind::tf32 (hi * hi, exact products) and ONE kind::f16 bf16 MMA of
//     K = 16 that adds both correction terms hi * lo + lo * hi; the dropped terms are < 2^-18 relative.
//   * tensor memory (512 columns): tiles x N accumulator columns (<= 384) + two A stages of 64 columns.
//   * a slot = 4 accumulator columns, owned by one task (a block's <= 4 components, or 4 background components) while
//     the walk is inside the block's rows.  When tasks end at a row the MMA thread commits, four epilogue warps read the
//     finished slots (tcgen05.ld), clear them (tcgen05.st), hand the accumulators back and store z; partial sums of
//     blocks that straddle two strips are added atomically (two contributions to a zeroed element: order independent).
// The background slots are additionally drained every <= 16 rows: the tensor core adds into its float32 accumulators
// with truncation, a bias that grows with the number of accumulation steps (see strips_ts_host.cu).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace pmd {

constexpr int kTSBStages = 3;
constexpr int kTSMaxRaw = 8;                 // raw movie stages (16 KB for float32) kept in flight, as many as fit
constexpr int kTSEpiWarps = 8, kTSEpiBatch = 8, kTSConvGroups = 2, kTSConvWarps = 4 * kTSConvGroups;
constexpr int kTSThreads = (kTSEpiWarps + kTSConvWarps + 3) * 32;   // + MMA warp + TMA warp + B loader warp
constexpr int kTSSmemBudget = 227 * 1024 - 2048;
constexpr int kTSAccCols = 384, kTSACols = 32;   // tensor memory: accumulators | 4 A stages of 16 pixels (hi 16 + pair 16 columns)
constexpr int kTSAStages = 4;

struct TSItem {                              // 12 ints (strips_ts_host.cu)
    int c0, nkc, row0, n_rows, b_chunk0, ev0, n_ev, part, slot_ptr0, n_drain, n_main, pad2;   // n_main: full-height items (first in the table)
};
struct TSEvent {
    int row, slot, col, ncw;
};

__device__ __forceinline__ uint32_t ts_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#ifdef PMD_TC_DEBUG
// debug build: a wait that does not complete within ~1 s reports which barrier is stuck and traps
__device__ __noinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 2000000000ll) {
            printf("stuck: block (%d,%d) thread %d barrier smem 0x%x parity %u\n", blockIdx.x, 0, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
#else
__device__ __forceinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra TS_DONE;\n\t"
        "bra TS_WAIT;\n\t"
        "TS_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
#endif
__device__ __forceinline__ bool ts_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void ts_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ts_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] += A[tmem] * B[smem descriptor]
__device__ __forceinline__ void ts_mma_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void ts_mma_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ uint32_t ts_pack_bf16(float lo_half, float hi_half) {   // lo_half -> bits [0,16), hi_half -> [16,32)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}
__device__ __forceinline__ uint4 ts_lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// Raw tile geometry per element type: a frame's 32 pixels are kSub rows of kRow bytes (one TMA box each), written with the
// TMA swizzle of that row length (address bits [4, 4 + log2(kRow / 16)) ^= address bits [7, ...)): conflict-free when every
// thread of a warp reads 16 bytes of its own row.
template <typename T>
struct TSRaw {
    static constexpr int kSub = sizeof(T) == 8 ? 2 : 1;
    static constexpr int kRow = 32 * (int)sizeof(T) / kSub;                 // 128, 64 or 32 bytes
    static constexpr int kSubTile = 128 * kRow;                              // bytes of one box
    static constexpr int kTile = kSub * kSubTile;
    static constexpr uint32_t kMask = kRow / 16 - 1;
    static constexpr int kPerPiece = 16 / (int)sizeof(T);                   // pixels per 16-byte piece
    // 16 consecutive pixels [16 h, 16 h + 16) of frame row m as float
    static __device__ __forceinline__ void load16(uint32_t tile, int m, int h, float* x) {
        constexpr int kPieces = 16 / kPerPiece;                              // 16-byte pieces per 16 pixels
#pragma unroll
        for (int q = 0; q < kPieces; ++q) {
            const int piece = h * kPieces + q;                               // piece index within the frame's 32 pixels
            const int sub = piece / (kRow / 16), pc = piece % (kRow / 16);
            uint32_t off = (uint32_t)(m * kRow + 16 * pc);
            off ^= ((off >> 7) & kMask) << 4;
            const uint4 v = ts_lds128(tile + sub * kSubTile + off);
            unpack(v, x + q * kPerPiece);
        }
    }
    static __device__ __forceinline__ void unpack(const uint4& v, float* x) {
        if constexpr (sizeof(T) == 4 && !std::is_integral<T>::value) {
            x[0] = __uint_as_float(v.x); x[1] = __uint_as_float(v.y); x[2] = __uint_as_float(v.z); x[3] = __uint_as_float(v.w);
        } else if constexpr (sizeof(T) == 4) {
            x[0] = (float)(int)v.x; x[1] = (float)(int)v.y; x[2] = (float)(int)v.z; x[3] = (float)(int)v.w;
        } else if constexpr (sizeof(T) == 8) {
            x[0] = (float)__hiloint2double((int)v.y, (int)v.x);
            x[1] = (float)__hiloint2double((int)v.w, (int)v.z);
        } else if constexpr (sizeof(T) == 2) {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if constexpr (std::is_signed<T>::value) {
                    x[2 * i] = (float)(int16_t)(w[i] & 0xFFFFu);
                    x[2 * i + 1] = (float)(int16_t)(w[i] >> 16);
                } else {
                    x[2 * i] = (float)(w[i] & 0xFFFFu);
                    x[2 * i + 1] = (float)(w[i] >> 16);
                }
            }
        } else {
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[4 * i] = (float)(w[i] & 0xFFu);
                x[4 * i + 1] = (float)((w[i] >> 8) & 0xFFu);
                x[4 * i + 2] = (float)((w[i] >> 16) & 0xFFu);
                x[4 * i + 3] = (float)(w[i] >> 24);
            }
        }
    }
};

template <typename T>
__global__ void __launch_bounds__(kTSThreads, 1)
project_ts_kernel(const __grid_constant__ CUtensorMap tm_movie, int64_t t, int d2, int64_t d, const TSItem* __restrict__ items,
                  const TSEvent* __restrict__ events, const unsigned char* __restrict__ bimg, const float* __restrict__ mean,
                  float* __restrict__ z, int64_t ldz, float* __restrict__ zbg, int64_t ldzbg, int64_t bg_stride, int n_cols_n,
                  int tiles, int n_raw, int movie_policy, int ablate, int n_items_total) {
    using Raw = TSRaw<T>;
    extern __shared__ __align__(1024) unsigned char tssm[];
    __shared__ __align__(8) uint64_t bar_rfull[kTSMaxRaw], bar_rempty[kTSMaxRaw], bar_afull[kTSAStages], bar_aempty[kTSAStages],
        bar_bfull[kTSBStages], bar_bempty[kTSBStages], bar_accfull, bar_accfree;
    __shared__ uint32_t tmem_base_s;
    const int N = n_cols_n;                                     // slot columns (UMMA N): 96, 128 or 192
    const int b_stage = 2 * N * 128;                            // bytes of one coefficient chunk (TF32 part + pair part)
    const uint32_t sbase = (ts_smem_u32(tssm) + 1023u) & ~1023u;
    const uint32_t sb_base = sbase;                             // B stages first (1024-byte aligned: N * 128 is a multiple of 1024)
    const uint32_t sr_base = sbase + kTSBStages * b_stage;      // raw movie stages
    const uint32_t sm_base = sr_base + n_raw * Raw::kTile;      // the 32 mean values of every raw stage's pixels (128 bytes each)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // grid: x = strip item (fastest: the CTAs that run together cover whole image rows), y = frame group
    // 1-D grid.  The n_main full-height items (one per strip) come first, strip index fastest: the CTAs that run together
    // walk the same image rows of the same frames, so a frame's row reaches DRAM as one burst of neighbouring 128-byte
    // requests.  The shorter left-over items follow (they would break that lock step) and fill the tail of the grid.
    int item_idx, fgroup;
    {
        const int n_main = items[0].n_main, n_fg = (int)((t + 128 * tiles - 1) / (128 * tiles)), bid = (int)blockIdx.x;
        if (bid < n_main * n_fg) {
            item_idx = bid % n_main;
            fgroup = bid / n_main;
        } else {
            const int r = bid - n_main * n_fg, n_extra = n_items_total - n_main;
            item_idx = n_main + r % n_extra;
            fgroup = r / n_extra;
        }
    }
    const TSItem it = items[item_idx];
    const int64_t f0 = (int64_t)fgroup * (128 * tiles);
    const int nft = (int)min((int64_t)tiles, (t - f0 + 127) / 128);   // frame tiles that hold at least one frame
    const int n_groups = it.n_rows * it.nkc;                          // (row, 32-pixel chunk) groups
    const int n_items = n_groups * nft;                               // (group, frame tile) work items

    if (tid == 0) {
        for (int s = 0; s < kTSMaxRaw; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ts_smem_u32(&bar_rfull[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;\n" ::"r"(ts_smem_u32(&bar_rempty[s])));
        }
        for (int s = 0; s < kTSAStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;\n" ::"r"(ts_smem_u32(&bar_afull[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ts_smem_u32(&bar_aempty[s])));
        }
        for (int s = 0; s < kTSBStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ts_smem_u32(&bar_bfull[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ts_smem_u32(&bar_bempty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ts_smem_u32(&bar_accfull)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(ts_smem_u32(&bar_accfree)), "r"(kTSEpiWarps * 32));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (tid < 32 * kTSMaxRaw) asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(sm_base + 4 * tid), "r"(0u) : "memory");   // (async-proxy writes follow the barrier below)
    constexpr uint32_t kCols = 512;   // the whole tensor memory of the SM
    constexpr int kMmaWarp = kTSEpiWarps + kTSConvWarps, kTmaWarp = kMmaWarp + 1, kBWarp = kMmaWarp + 2;
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(ts_smem_u32(&tmem_base_s)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    if (warp == kTmaWarp && lane == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_movie) : "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem_d = tmem_base_s;

    if (warp < kTSEpiWarps) {
        // ================================ epilogue warps ================================
        // warp w reads / writes the tensor-memory lanes 32 (w & 3) .. + 31 (= frames of a tile) of the frame tiles
        // 2 (w >> 2) and 2 (w >> 2) + 1
        const int quad = warp & 3, ft0 = 2 * (warp >> 2);
        const uint32_t lane_base = tmem_d + ((uint32_t)(32 * quad) << 16);
        for (int c = (warp >> 2) * (kTSAccCols / 2); c < ((warp >> 2) + 1) * (kTSAccCols / 2); c += 16) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(
                             lane_base + c),
                         "r"(0u)
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
        ts_mbar_arrive(ts_smem_u32(&bar_accfree));             // completion 0: accumulators are zero
        const int4* ev = reinterpret_cast<const int4*>(events + it.ev0);   // (row, slot, first column, n comps | kind << 8)
        int e = (ablate & 2) ? it.n_ev : 0, k = 0;              // k = index of the drain row
        bool row_open = false;                                  // the accumulators of drain row k are already ours
        while (e < it.n_ev) {
            // Every lane fetches one event of the window [e, e + 32) BEFORE the warp waits for the accumulators (the
            // list lives in global memory: its latency must not sit between the MMA warp's commit and our release).
            // Events are sorted by row: the drain row's events are a prefix of the window; lane 31 is look-ahead only.
            int4 my = make_int4(-1, 0, 0, 0);
            if (e + lane < it.n_ev) my = __ldg(ev + e + lane);
            const int row = __shfl_sync(0xffffffffu, my.x, 0);
            const unsigned same = __ballot_sync(0xffffffffu, my.x == row);
            const bool last_chunk = same != 0xffffffffu;         // the row's events end inside this window
            const int cnt = last_chunk ? __ffs(~same) - 1 : 31;
            if (!row_open) {
                ts_mbar_wait(ts_smem_u32(&bar_accfull), k & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                row_open = true;
            }
            // The MMA stream stands still until the accumulators are handed back, so a pass first moves up to kTSEpiBatch
            // finished slots (this warp's two frame tiles of each) into registers and clears them -- back-to-back tensor-
            // memory instructions, one wait each -- releases the accumulators if these were the row's last events, and
            // only then forms addresses and issues the global stores / reductions.
            for (int eb = 0; eb < cnt; eb += kTSEpiBatch) {
                uint32_t v[kTSEpiBatch][2][4];
#pragma unroll
                for (int j = 0; j < kTSEpiBatch; ++j) {
                    if (eb + j < cnt) {
                        const int slot = __shfl_sync(0xffffffffu, my.y, min(eb + j, 31));
                        const uint32_t ta = lane_base + 4 * slot + N * ft0;
#pragma unroll
                        for (int f = 0; f < 2; ++f)
                            if (ft0 + f < nft)
                                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                                             : "=r"(v[j][f][0]), "=r"(v[j][f][1]), "=r"(v[j][f][2]), "=r"(v[j][f][3])
                                             : "r"(ta + N * f));
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int j = 0; j < kTSEpiBatch; ++j) {
                    if (eb + j < cnt) {
                        const int slot = __shfl_sync(0xffffffffu, my.y, min(eb + j, 31));
                        const uint32_t ta = lane_base + 4 * slot + N * ft0;
#pragma unroll
                        for (int f = 0; f < 2; ++f)
                            if (ft0 + f < nft)
                                asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(ta + N * f), "r"(0u)
                                             : "memory");
                    }
                }
                if (last_chunk && eb + kTSEpiBatch >= cnt) {     // last pass of this row: hand the accumulators back
                    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                    ts_mbar_arrive(ts_smem_u32(&bar_accfree));
                }
#pragma unroll
                for (int j = 0; j < kTSEpiBatch; ++j) {
                    if (eb + j < cnt) {
                        const int colv = __shfl_sync(0xffffffffu, my.z, min(eb + j, 31));
                        const int ncw = __shfl_sync(0xffffffffu, my.w, min(eb + j, 31));
                        const int nc = ncw & 0xFF, kind = ncw >> 8;
                        float* zo;
                        int64_t ldo;
                        if (kind != 1) {
                            zo = z + (int64_t)colv * ldz;
                            ldo = ldz;
                        } else {
                            zo = zbg + (int64_t)it.part * bg_stride + (int64_t)colv * ldzbg;
                            ldo = ldzbg;
                        }
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const int64_t fr = f0 + 128 * (ft0 + f) + 32 * quad + lane;
                            if (ft0 + f < nft && fr < t && !(ablate & 16)) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    if (c < nc) {
                                        float* o = zo + (int64_t)c * ldo + fr;
                                        const float val = __uint_as_float(v[j][f][c]);
                                        // kind 1: only this thread ever touches the element of the strip's partial buffer, and
                                        // its adds reach the element in program order (deterministic); a reduction instead of
                                        // load-add-store keeps dependent L2 round trips out of the drain.
                                        // kind 2: block shared by two strips (z zeroed by the caller)
                                        if (kind == 0) *o = val;
                                        else atomicAdd(o, val);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            e += cnt;
            if (last_chunk) {
                ++k;
                row_open = false;
            }
        }
    } else if (warp < kMmaWarp) {
        // ================================ converters ================================
        // group j (4 warps = 128 threads, thread = frame) handles the work items i = j, j + 2, ... : raw tile -> centre ->
        // hi / bf16 pair in registers (before it waits for tensor memory), then one tcgen05.st pair per 16-pixel half
        // into A stages 2 j and 2 j + 1, each handed to the MMA warp as soon as it is written
        const int cw = warp - kTSEpiWarps, j = cw >> 2;
        const int m = 32 * (warp & 3) + lane;                                    // frame within the tile = tensor-memory lane
        const uint32_t ta0 = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + kTSAccCols + kTSACols * 2 * j;
        const uint32_t afull = ts_smem_u32(&bar_afull[2 * j]), aempty = ts_smem_u32(&bar_aempty[2 * j]);
        int rs = j % n_raw;                                                      // raw stage of item i, its use count
        uint32_t ruse = (uint32_t)(j / n_raw);
        for (int i = j, n = 0; i < n_items; i += kTSConvGroups, ++n) {
            if (!(ablate & 8)) ts_mbar_wait(ts_smem_u32(&bar_rfull[rs]), ruse & 1);
            const uint32_t tile = sr_base + rs * Raw::kTile;
            uint32_t hi[32], pr[32];
            if (ablate & 4) {
#pragma unroll
                for (int q = 0; q < 32; ++q) hi[q] = pr[q] = (uint32_t)(n + q);
            }
#pragma unroll
            for (int h = 0; h < 2 && !(ablate & 4); ++h) {
                float x[16];
                Raw::load16(tile, m, h, x);
                if (mean) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                                // broadcast reads: one wavefront each
                        const uint4 mu = ts_lds128(sm_base + rs * 128 + 64 * h + 16 * q);
                        x[4 * q] -= __uint_as_float(mu.x); x[4 * q + 1] -= __uint_as_float(mu.y);
                        x[4 * q + 2] -= __uint_as_float(mu.z); x[4 * q + 3] -= __uint_as_float(mu.w);
                    }
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    hi[16 * h + q] = __float_as_uint(x[q]) & 0xFFFFE000u;
                    pr[16 * h + q] = ts_pack_bf16(__uint_as_float(hi[16 * h + q]), x[q] - __uint_as_float(hi[16 * h + q]));
                }
            }
            if (!(ablate & 8)) ts_mbar_arrive(ts_smem_u32(&bar_rempty[rs]));     // the raw stage has been read (values are in registers)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (n >= 1) {
                    ts_mbar_wait(aempty + 8 * h, (n - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                }
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(
                        ta0 + kTSACols * h),
                    "r"(hi[16 * h + 0]), "r"(hi[16 * h + 1]), "r"(hi[16 * h + 2]), "r"(hi[16 * h + 3]), "r"(hi[16 * h + 4]), "r"(hi[16 * h + 5]),
                    "r"(hi[16 * h + 6]), "r"(hi[16 * h + 7]), "r"(hi[16 * h + 8]), "r"(hi[16 * h + 9]), "r"(hi[16 * h + 10]), "r"(hi[16 * h + 11]),
                    "r"(hi[16 * h + 12]), "r"(hi[16 * h + 13]), "r"(hi[16 * h + 14]), "r"(hi[16 * h + 15])
                    : "memory");
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(
                        ta0 + kTSACols * h + 16),
                    "r"(pr[16 * h + 0]), "r"(pr[16 * h + 1]), "r"(pr[16 * h + 2]), "r"(pr[16 * h + 3]), "r"(pr[16 * h + 4]), "r"(pr[16 * h + 5]),
                    "r"(pr[16 * h + 6]), "r"(pr[16 * h + 7]), "r"(pr[16 * h + 8]), "r"(pr[16 * h + 9]), "r"(pr[16 * h + 10]), "r"(pr[16 * h + 11]),
                    "r"(pr[16 * h + 12]), "r"(pr[16 * h + 13]), "r"(pr[16 * h + 14]), "r"(pr[16 * h + 15])
                    : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                ts_mbar_arrive(afull + 8 * h);
            }
            rs += kTSConvGroups;
            while (rs >= n_raw) {
                rs -= n_raw;
                ++ruse;
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================ MMA issuer ================================
        // The whole warp runs the (warp-uniform) control flow so that addresses and descriptors stay in uniform
        // registers; one elected lane issues the tcgen05 instructions of a work item and its commits.
        // D f32, A K-major from tensor memory, B K-major SWIZZLE_128B from shared memory, M = 128, N = n_cols_n
        const uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        // K-major SWIZZLE_128B descriptor: low word = start address >> 4, high word = SBO 1024 >> 4 | version | layout type
        constexpr uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const TSEvent* ev = events + it.ev0;
        int e = 0, k = 0;
        // rows of the events [e_base, e_base + 32), one per lane (the list is sorted by row): the next drain row comes
        // from registers instead of a chain of dependent global loads
        int e_base = 0;
        int myrow = lane < it.n_ev ? ev[lane].row : 0x7fffffff;
        const uint32_t afull0 = ts_smem_u32(&bar_afull[0]), aempty0 = ts_smem_u32(&bar_aempty[0]);
        const uint32_t bfull0 = ts_smem_u32(&bar_bfull[0]), bempty0 = ts_smem_u32(&bar_bempty[0]);
        const uint32_t accfull = ts_smem_u32(&bar_accfull), accfree = ts_smem_u32(&bar_accfree);
        const bool leader = ts_elect_one();
        const uint32_t b_tf0 = sb_base >> 4;
        ts_mbar_wait(accfree, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        int kc = 0, row = it.row0, bs = 0;
        uint32_t bpar = 0;
        int next_ev_row = it.n_ev > 0 ? __shfl_sync(0xffffffffu, myrow, 0) : -1;
        int i = 0;
        for (int g = 0; g < n_groups; ++g) {
            ts_mbar_wait(bfull0 + 8 * bs, bpar);
            const uint32_t b_tf = b_tf0 + bs * (b_stage >> 4), b_bf = b_tf + ((N * 128) >> 4);
            for (int ft = 0; ft < nft; ++ft, ++i) {
                const uint32_t dcol = tmem_d + N * ft;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int as = (2 * i + h) & (kTSAStages - 1);              // A stage of this 16-pixel half, its use (2 i + h) / 4
                    ts_mbar_wait(afull0 + 8 * as, ((2 * i + h) >> 2) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    if (leader) {
                        const uint32_t a_hi = tmem_d + kTSAccCols + kTSACols * as, a_pr = a_hi + 16;
#pragma unroll
                        for (int ks = 0; ks < 2 && !(ablate & 1); ++ks) {
                            ts_mma_tf32(dcol, a_hi + 8 * ks, desc_hi | (b_tf + 2 * (2 * h + ks)), idesc_tf32);
                            ts_mma_bf16(dcol, a_pr + 8 * ks, desc_hi | (b_bf + 2 * (2 * h + ks)), idesc_bf16);
                        }
                        ts_commit(aempty0 + 8 * as);
                        if (h == 1 && ft == nft - 1) ts_commit(bempty0 + 8 * bs);
                    }
                }
            }
            if (++bs == kTSBStages) {
                bs = 0;
                bpar ^= 1;
            }
            if (kc == it.nkc - 1 && row == next_ev_row && !(ablate & 2)) {
                // tasks end at this row: let the epilogue warps drain and clear their slots
                for (;;) {                                        // skip the events of this row
                    e += __popc(__ballot_sync(0xffffffffu, myrow == row));
                    if (e - e_base < 32) break;
                    e_base = e;
                    myrow = e_base + lane < it.n_ev ? ev[e_base + lane].row : 0x7fffffff;
                }
                next_ev_row = e < it.n_ev ? __shfl_sync(0xffffffffu, myrow, e - e_base) : -1;
                if (leader) ts_commit(accfull);
                ++k;
                ts_mbar_wait(accfree, k & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            }
            if (++kc == it.nkc) {
                kc = 0;
                ++row;
            }
        }
    } else if (warp == kTmaWarp) {
        // ================================ raw movie tiles (one thread) ================================
        if (lane == 0 && !(ablate & 8)) {
            uint64_t pol_stream;                                                  // the movie is read once: do not let it push the
            if (movie_policy == 0)                                                // coefficient images out of L2
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol_stream));
            else
                asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;\n" : "=l"(pol_stream));
            int rs = 0;
            uint32_t ruse = 0;
            int i = 0;
            for (int g = 0; g < n_groups; ++g) {
                const int row = it.row0 + g / it.nkc, kc = g % it.nkc;
                const int x = row * d2 + it.c0 + 32 * kc;
                const int mean_bytes = mean ? (int)min((int64_t)128, 4 * (d - x)) : 0;   // the last chunk of the frame may be short
                for (int ft = 0; ft < nft; ++ft, ++i) {
                    if (ruse >= 1) ts_mbar_wait(ts_smem_u32(&bar_rempty[rs]), (ruse - 1) & 1);
                    const uint32_t bar = ts_smem_u32(&bar_rfull[rs]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)(Raw::kTile + mean_bytes))
                                 : "memory");
                    const int y = (int)(f0 + 128 * ft);
#pragma unroll
                    for (int sub = 0; sub < Raw::kSub; ++sub)
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(
                                sr_base + rs * Raw::kTile + sub * Raw::kSubTile),
                            "l"(&tm_movie), "r"(x + sub * (32 / Raw::kSub)), "r"(y), "r"(bar), "l"(pol_stream)
                            : "memory");
                    if (mean_bytes > 0)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                         sm_base + rs * 128),
                                     "l"(mean + x), "r"((uint32_t)mean_bytes), "r"(bar)
                                     : "memory");
                    if (++rs == n_raw) {
                        rs = 0;
                        ++ruse;
                    }
                }
            }
        }
    } else if (warp == kBWarp) {
        // ================================ coefficient chunks (one thread) ================================
        if (lane == 0) {
            const unsigned char* src = bimg + (int64_t)it.b_chunk0 * b_stage;
            const uint32_t bfull0 = ts_smem_u32(&bar_bfull[0]), bempty0 = ts_smem_u32(&bar_bempty[0]);
            uint64_t pol_keep;                                                    // shared by every frame group of the strip
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol_keep));
            for (int g = 0; g < n_groups; ++g) {
                const int bs = g % kTSBStages;
                if (g >= kTSBStages) ts_mbar_wait(bempty0 + 8 * bs, ((g / kTSBStages) - 1) & 1);
                const uint32_t bar = bfull0 + 8 * bs;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)b_stage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                                 sb_base + bs * b_stage),
                             "l"(src + (int64_t)g * b_stage), "r"((uint32_t)b_stage), "r"(bar), "l"(pol_keep)
                             : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(kCols));
}

// Coefficient images: one CTA per (item, row); chunk kc of the row is the shared-memory image (2 N 128 bytes)
//   part 0: [N slot columns][32 pixels] float32, TF32-exact hi values of u / std, K-major SWIZZLE_128B
//   part 1: same shape, per pixel the bf16 pair (bf16(lo) in bits [0,16), bf16(hi) in [16,32))
__global__ void __launch_bounds__(256)
pack_strips_ts_kernel(const TSItem* __restrict__ items, const int32_t* __restrict__ item_of_row, const int32_t* __restrict__ slot_ptr,
                      const int32_t* __restrict__ tasks, const float* __restrict__ uvals, int64_t bpix, const float* __restrict__ bg,
                      const float* __restrict__ inv_std, int64_t d, int64_t d2, int N, unsigned char* __restrict__ bimg) {
    __shared__ int s_task[48];
    const int ii = item_of_row[2 * blockIdx.x], rr = item_of_row[2 * blockIdx.x + 1];
    const TSItem it = items[ii];
    const int row = it.row0 + rr;
    const int n_slots = N / 4;
    if ((int)threadIdx.x < n_slots) {
        int found = -1;
        const int a = slot_ptr[it.slot_ptr0 + threadIdx.x], b = slot_ptr[it.slot_ptr0 + threadIdx.x + 1];
        for (int i = a; i < b; ++i) {
            const int by = tasks[8 * i], h = tasks[8 * i + 2];
            if (row >= by && row < by + h) found = i;
        }
        s_task[threadIdx.x] = found;
    }
    __syncthreads();
    const int64_t b_stage = 2ll * N * 128;
    unsigned char* out = bimg + ((int64_t)it.b_chunk0 + (int64_t)rr * it.nkc) * b_stage;
    const int pieces = it.nkc * N * 8;     // 16-byte pieces (4 pixels of one column) of one part
    for (int p = threadIdx.x; p < pieces; p += blockDim.x) {
        const int c = p & 7, n = (p >> 3) % N, kc = p / (8 * N);
        const int ti = s_task[n >> 2], comp = n & 3;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (ti >= 0) {
            const int32_t* tk = tasks + 8 * ti;
            if (comp < tk[5]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xs = 32 * kc + 4 * c + j;            // column within the strip
                    const int x = xs - tk[1];                      // column within the task's block
                    const int64_t pix = (int64_t)row * d2 + it.c0 + xs;
                    if (x >= 0 && x < tk[3] && it.c0 + xs < d2) {
                        const float u = tk[6] != 1 ? uvals[(int64_t)(tk[4] + comp) * bpix + (int64_t)(row - tk[0]) * tk[3] + x]
                                                   : bg[(int64_t)(tk[4] + comp) * d + pix];
                        v[j] = inv_std ? u * inv_std[pix] : u;
                    }
                }
            }
        }
        float hi[4];
        uint32_t pr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
            pr[j] = ts_pack_bf16(v[j] - hi[j], hi[j]);
        }
        const int64_t off = (int64_t)kc * b_stage + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4);
        *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(out + off + (int64_t)N * 128) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
    }
}

typedef CUresult (*TSEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TSEncodeFn ts_encode_fn() {
    static TSEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (TSEncodeFn)p;
    }();
    return fn;
}

template <typename T>
static int launch_project_ts(const void* movie, int64_t t, int64_t d2, int64_t d, const int32_t* items, int64_t n_items,
                             const int32_t* events, const void* bimg, const float* mean, float* z, int64_t ldz, float* zbg,
                             int64_t ldzbg, int64_t bg_stride, int n, cudaStream_t st, const char* fn) {
    using Raw = TSRaw<T>;
    TSEncodeFn enc = ts_encode_fn();
    if (!enc) {
        set_error(std::string(fn) + ": cuTensorMapEncodeTiled is not available");
        return -2;
    }
    CUtensorMapDataType dt = sizeof(T) == 8   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                             : sizeof(T) == 4 ? (std::is_integral<T>::value ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)
                             : sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                              : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUtensorMapSwizzle sw = Raw::kRow == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : Raw::kRow == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)t};
    cuuint64_t gstr[1] = {(cuuint64_t)d * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)(32 / Raw::kSub), 128u};
    cuuint32_t estr[2] = {1, 1};
    // 128-byte promotion: a box row is one 128-byte line; with 256 bytes the neighbour strip's half was fetched, evicted (evict_first)
    // and fetched again (measured at C2: 30.6 GB of DRAM reads and 5.68 ms against 27.4 GB and 5.32 ms)
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    int movie_policy = 0, ablate = 0;   // ablate: profiling aid of -DPMD_TUNE builds (results are wrong when set)
#ifdef PMD_TUNE
    if (const char* e = getenv("PMD_TS_ABLATE")) ablate = atoi(e);   // development builds only: access-shape experiments
    if (const char* e = getenv("PMD_TS_PROMO")) promo = atoi(e) == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : atoi(e) == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : atoi(e) == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* e = getenv("PMD_TS_POLICY")) movie_policy = atoi(e);
#endif
    CUresult r = enc(&tm, dt, 2, const_cast<void*>(movie), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(std::string(fn) + ": cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return -3;
    }
    if (((uint64_t)d * sizeof(T)) % 16 != 0 || ((uint64_t)d2 * sizeof(T)) % 16 != 0)
        return fail_arg(fn, "frame and image-row pitch must be multiples of 16 bytes (TMA box origins)");
    const int tiles = kTSAccCols / n;
    const int b_stage = 2 * n * 128;
    int n_raw = (kTSSmemBudget - kTSBStages * b_stage - 128 * kTSMaxRaw) / Raw::kTile;
    if (n_raw > kTSMaxRaw) n_raw = kTSMaxRaw;
    if (n_raw < 2) return fail_arg(fn, "shared memory too small for this element type");
    const int smem = kTSBStages * b_stage + n_raw * Raw::kTile + 128 * kTSMaxRaw + 1024;
    auto k = project_ts_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        set_error(std::string(fn) + ": " + cudaGetErrorString(e));
        return (int)e;
    }
    const int64_t fgroups = (t + 128 * tiles - 1) / (128 * tiles);
    k<<<dim3((unsigned)(n_items * fgroups)), kTSThreads, smem, st>>>(tm, t, (int)d2, d, (const TSItem*)items, (const TSEvent*)events,
                                                                            (const unsigned char*)bimg, mean, z, ldz, zbg, ldzbg, bg_stride, n,
                                                                            tiles, n_raw, movie_policy, ablate, (int)n_items);
    return check_launch(fn);
}

}  // namespace pmd

extern "C" int pmd_pack_strips_ts(const int32_t* items, const int32_t* item_of_row, int64_t n_rows_total, const int32_t* slot_ptr,
                                  const int32_t* tasks, const float* uvals, int64_t bpix, const float* bg, const float* inv_std,
                                  int64_t d, int64_t d2, int64_t n, void* bimg, void* stream) {
    const char* fn = "pmd_pack_strips_ts";
    PMD_REQUIRE(items && item_of_row && slot_ptr && tasks && bimg, fn, "null pointer");
    PMD_REQUIRE(n_rows_total > 0 && bpix > 0 && d > 0 && d2 > 0, fn, "bad size");
    PMD_REQUIRE(n == 96 || n == 128 || n == 192, fn, "N must be 96, 128 or 192");
    PMD_REQUIRE(((uintptr_t)bimg & 15) == 0, fn, "image must be 16-byte aligned");
    pmd::pack_strips_ts_kernel<<<(unsigned)n_rows_total, 256, 0, (cudaStream_t)stream>>>(
        (const pmd::TSItem*)items, item_of_row, slot_ptr, tasks, uvals, bpix, bg, inv_std, d, d2, (int)n, (unsigned char*)bimg);
    return pmd::check_launch(fn);
}

extern "C" int pmd_project_stream_ts(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* items,
                                     int64_t n_items, const int32_t* events, const void* bimg, int64_t n, const float* mean, float* z,
                                     int64_t ldz, float* zbg, int64_t ldzbg, int64_t bg_stride, void* stream) {
    const char* fn = "pmd_project_stream_ts";
    PMD_REQUIRE(movie && items && events && bimg && z && zbg, fn, "null pointer");
    PMD_REQUIRE(t > 0 && n_items > 0 && ldz >= t && ldzbg >= t && d2 > 0 && d >= d2, fn, "bad size");
    PMD_REQUIRE(n == 96 || n == 128 || n == 192, fn, "N must be 96, 128 or 192");
    PMD_REQUIRE((d2 & 3) == 0 && (d & 3) == 0, fn, "row length must be a multiple of 4 pixels");
    PMD_REQUIRE(d < (1ll << 31) && t < (1ll << 31), fn, "movie too large for 32-bit TMA coordinates");
    PMD_REQUIRE(((uintptr_t)movie & 15) == 0 && ((uintptr_t)bimg & 15) == 0, fn, "movie and image must be 16-byte aligned");
    PMD_REQUIRE(!mean || ((uintptr_t)mean & 15) == 0, fn, "mean must be 16-byte aligned");
    PMD_REQUIRE(n_items * ((t + 255) / 256) <= 0x7FFFFFFF, fn, "too many strip items");
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        return pmd::launch_project_ts<scalar_t>(movie, t, d2, d, items, n_items, events, bimg, mean, z, ldz, zbg, ldzbg, bg_stride, (int)n,
                                                st, fn);
    });
    return 0;
}
