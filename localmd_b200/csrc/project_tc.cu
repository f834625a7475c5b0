// K7 on the 5th-generation tensor cores:  z = U^T ((Y - mean) / std)  over the whole movie in ONE streaming pass
// (pmd_loader.py:316-346, 392-414: v_projection / v_projection_routine), local and background columns together.
//
// A CTA owns a column strip of the field of view (host tables: strips_tc_host.cu), a range of image rows and
// kPTTiles x 128 frames, and walks the strip one image row at a time.  Per row and 32-pixel chunk:
//   D[128 frames x 128 slot columns] += A[128 frames x 32 pixels] * B[32 pixels x 128 slot columns]   (per frame tile)
//   * A: the movie is frame-major, so the 32 pixels of a frame are 128 contiguous bytes = one row of the canonical
//     K-major SWIZZLE_128B operand.  Sixteen producer warps, four per frame tile, load them (128-bit loads, all pieces
//     of the next chunk issued together one chunk ahead), centre / scale, and write TWO operand tiles: hi = x with the
//     low 13 mantissa bits cleared (exact in TF32) and a bf16 pair tile (bf16(hi), bf16(x - hi)) per pixel.  A
//     prefetch warp asks L2 for the next strip row of every frame as one contiguous burst.
//   * B: the coefficient image of the strip row, prebuilt once per decomposition by pmd_pack_strips_tc in exactly
//     the shared-memory image (swizzled, TF32 hi part + bf16 pair part (bf16(lo), bf16(hi))); one thread fetches each
//     32 KB chunk with a single bulk asynchronous copy (cp.async.bulk, mbarrier complete_tx).
//   * one thread issues, per 8 pixels, tcgen05.mma.kind::tf32 (hi * hi, exact products) and ONE kind::f16 bf16 MMA of
//     K = 16 that adds both correction terms  hi * lo + lo * hi;  the dropped terms are < 2^-18 relative.
//     Accumulators live in tensor memory: 4 frame tiles x 128 columns = all 512 columns.
//   * a slot = 4 accumulator columns, owned by one task (a block's <= 4 components, or 4 background components) while
//     the walk is inside the block's rows.  When tasks end at a row, the MMA thread commits, four epilogue warps read
//     the finished slots (tcgen05.ld), clear them (tcgen05.st) and hand the accumulators back, then store z.
// Every movie element is read once per strip that contains it (strips overlap by half a block only).
// The background slots are additionally drained every <= 16 rows: the tensor core adds into its float32 accumulators
// with truncation, a bias that grows with the number of accumulation steps (see strips_tc_host.cu).
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace pmd {

constexpr int kPTTiles = 4;                  // 128-frame accumulator tiles per CTA
constexpr int kPTN = 128;                    // slot columns (UMMA N) = 32 slots x 4
constexpr int kPTSlotCols = 4;
constexpr int kPTAStages = kPTTiles;      // one stage per frame tile: its producer pair and the MMA thread alternate on it
constexpr int kPTBStages = 3;             // a 32 KB coefficient chunk takes longer to arrive than its MMAs take to run
constexpr int kPTBPrefetch = 8;           // L2 prefetch distance of the coefficient chunks (chunks)
constexpr int kPTATile = 128 * 128;          // bytes of one A operand tile (128 frames x 32 pixels x 4 B)
constexpr int kPTAStage = 2 * kPTATile;      // tf32 tile + bf16 pair tile
constexpr int kPTBPart = kPTN * 128;         // bytes of one B part (128 columns x 32 pixels x 4 B)
constexpr int kPTBStage = 2 * kPTBPart;      // 32 KB
constexpr int kPTEpiWarps = 4, kPTProdWarps = 16;
constexpr int kPTProducers = kPTProdWarps * 32;
constexpr int kPTThreads = (kPTEpiWarps + kPTProdWarps + 3) * 32;   // + MMA warp + B loader warp + L2 prefetch warp
constexpr int kPTSmem = kPTAStages * kPTAStage + kPTBStages * kPTBStage + 1024;

struct PTItem {                              // 12 ints (strips_tc_host.cu)
    int c0, w8, row0, n_rows, b_chunk0, nkc, ev0, n_ev, part, slot_ptr0, pad0, pad1;
};
struct PTEvent {
    int row, slot, col, ncw;
};

__device__ __forceinline__ uint32_t pt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#ifdef PMD_TC_DEBUG
// debug build: a wait that does not complete within ~1 s reports which barrier is stuck and traps
__device__ __noinline__ void pt_mbar_wait(uint32_t bar, uint32_t parity) {
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
            printf("stuck: block (%d,%d) thread %d barrier smem 0x%x parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
#else
__device__ __forceinline__ void pt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "PT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra PT_DONE;\n\t"
        "bra PT_WAIT;\n\t"
        "PT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
#endif
__device__ __forceinline__ bool pt_elect_one() {
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
__device__ __forceinline__ void pt_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pt_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// K-major SWIZZLE_128B operand descriptor: 128-byte rows, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t pt_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void pt_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void pt_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
        : "memory");
}

// four consecutive pixels of one frame as float
template <typename T>
struct Px4;
template <>
struct Px4<float> {
    float4 v;
    __device__ __forceinline__ void load(const float* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    }
    __device__ __forceinline__ float4 get() const { return v; }
};
template <>
struct Px4<uint16_t> {
    uint2 v;
    __device__ __forceinline__ void load(const uint16_t* p) { v = __ldg(reinterpret_cast<const uint2*>(p)); }
    __device__ __forceinline__ float4 get() const {
        return make_float4((float)(v.x & 0xFFFFu), (float)(v.x >> 16), (float)(v.y & 0xFFFFu), (float)(v.y >> 16));
    }
};
template <>
struct Px4<int16_t> {
    uint2 v;
    __device__ __forceinline__ void load(const int16_t* p) { v = __ldg(reinterpret_cast<const uint2*>(p)); }
    __device__ __forceinline__ float4 get() const {
        return make_float4((float)(int16_t)(v.x & 0xFFFFu), (float)(int16_t)(v.x >> 16), (float)(int16_t)(v.y & 0xFFFFu),
                           (float)(int16_t)(v.y >> 16));
    }
};
template <>
struct Px4<uint8_t> {
    uint32_t v;
    __device__ __forceinline__ void load(const uint8_t* p) { v = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    __device__ __forceinline__ float4 get() const {
        return make_float4((float)(v & 0xFFu), (float)((v >> 8) & 0xFFu), (float)((v >> 16) & 0xFFu), (float)(v >> 24));
    }
};
template <>
struct Px4<int32_t> {
    int4 v;
    __device__ __forceinline__ void load(const int32_t* p) { v = __ldg(reinterpret_cast<const int4*>(p)); }
    __device__ __forceinline__ float4 get() const { return make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w); }
};
template <>
struct Px4<double> {
    double2 a, b;
    __device__ __forceinline__ void load(const double* p) {
        a = __ldg(reinterpret_cast<const double2*>(p));
        b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    }
    __device__ __forceinline__ float4 get() const { return make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y); }
};

__device__ __forceinline__ uint32_t pt_pack_bf16(float lo_half, float hi_half) {   // lo_half -> bits [0,16), hi_half -> [16,32)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}

template <typename T>
__global__ void __launch_bounds__(kPTThreads, 1)
project_tc_kernel(const T* __restrict__ movie, int64_t t, int64_t d2, int64_t d, const PTItem* __restrict__ items,
                  const PTEvent* __restrict__ events, const unsigned char* __restrict__ bimg, const float* __restrict__ mean,
                  const float* __restrict__ inv_std, float* __restrict__ z, int64_t ldz, float* __restrict__ zbg, int64_t ldzbg,
                  int64_t bg_stride, int ablate, int pf_rows) {
    extern __shared__ __align__(1024) unsigned char ptsm[];
    __shared__ __align__(8) uint64_t bar_afull[kPTAStages], bar_aempty[kPTAStages], bar_bfull[kPTBStages], bar_bempty[kPTBStages],
        bar_accfull, bar_accfree;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int s_rows_issued;
    const uint32_t sbase = (pt_smem_u32(ptsm) + 1023u) & ~1023u;
    const uint32_t sb_base = sbase + kPTAStages * kPTAStage;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // grid: x = frame tile (fastest), y = item: the CTAs of one strip run together and share its coefficient images in L2
    const PTItem it = items[blockIdx.y];
    const int64_t f0 = (int64_t)blockIdx.x * (128 * kPTTiles);
    const int nft = (int)min((int64_t)kPTTiles, (t - f0 + 127) / 128);   // frame tiles that hold at least one frame
    const int n_groups = it.n_rows * it.nkc;                              // (row, 32-pixel chunk) groups

    if (tid == 0) {
        s_rows_issued = 0;
        for (int s = 0; s < kPTAStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(pt_smem_u32(&bar_afull[s])), "r"(kPTProducers / kPTTiles));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(pt_smem_u32(&bar_aempty[s])));
        }
        for (int s = 0; s < kPTBStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(pt_smem_u32(&bar_bfull[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(pt_smem_u32(&bar_bempty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(pt_smem_u32(&bar_accfull)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(pt_smem_u32(&bar_accfree)), "r"(kPTEpiWarps * 32));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    constexpr uint32_t kCols = kPTN * kPTTiles;   // 512: the whole tensor memory of the SM
    if (warp == kPTEpiWarps + kPTProdWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(pt_smem_u32(&tmem_base_s)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem_d = tmem_base_s;

    if (warp < kPTEpiWarps) {
        // ================================ epilogue warps ================================
        // warp q reads / writes the tensor-memory lanes 32 q .. 32 q + 31 (= frames of a tile)
        const uint32_t lane_base = tmem_d + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < (int)kCols; c += 16) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(
                             lane_base + c),
                         "r"(0u)
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
        pt_mbar_arrive(pt_smem_u32(&bar_accfree));             // completion 0: accumulators are zero
        const PTEvent* ev = events + it.ev0;
        int e = 0, k = 0;                                       // k = index of the drain row
        while (e < it.n_ev) {
            const int row = ev[e].row;
            int e1 = e + 1;
            while (e1 < it.n_ev && ev[e1].row == row) ++e1;
            pt_mbar_wait(pt_smem_u32(&bar_accfull), k & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            for (int eb = e; eb < e1; eb += 2) {                 // two events (x 4 frame tiles x 4 columns) per batch
                uint32_t v[2][kPTTiles][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (eb + j < e1) {
                        const uint32_t ta = lane_base + kPTSlotCols * ev[eb + j].slot;
#pragma unroll
                        for (int ft = 0; ft < kPTTiles; ++ft)
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                                         : "=r"(v[j][ft][0]), "=r"(v[j][ft][1]), "=r"(v[j][ft][2]), "=r"(v[j][ft][3])
                                         : "r"(ta + kPTN * ft));
                    }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (eb + j < e1) {
                        const uint32_t ta = lane_base + kPTSlotCols * ev[eb + j].slot;
#pragma unroll
                        for (int ft = 0; ft < kPTTiles; ++ft)
                            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(ta + kPTN * ft), "r"(0u)
                                         : "memory");
                    }
                }
                if (eb + 2 >= e1) {                              // last batch of this row: hand the accumulators back
                    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                    pt_mbar_arrive(pt_smem_u32(&bar_accfree));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (eb + j < e1) {
                        const PTEvent evj = ev[eb + j];
                        const int nc = evj.ncw & 0xFF, kind = evj.ncw >> 8;
                        float* zo;
                        int64_t ldo;
                        if (kind == 0) {
                            zo = z + (int64_t)evj.col * ldz;
                            ldo = ldz;
                        } else {
                            zo = zbg + (int64_t)it.part * bg_stride + (int64_t)evj.col * ldzbg;
                            ldo = ldzbg;
                        }
#pragma unroll
                        for (int ft = 0; ft < kPTTiles; ++ft) {
                            const int64_t f = f0 + 128 * ft + 32 * warp + lane;
                            if (f < t) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    if (c < nc) {
                                        float* o = zo + (int64_t)c * ldo + f;
                                        // background partial sums are added (this thread owns the element: no race)
                                        *o = kind == 0 ? __uint_as_float(v[j][ft][c]) : *o + __uint_as_float(v[j][ft][c]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            e = e1;
            ++k;
        }
    } else if (warp < kPTEpiWarps + kPTProdWarps) {
        // ================================ A producers ================================
        // Four warps (128 threads) own one frame tile: they fill that tile's stage once per (row, 32-pixel chunk) group.
        // A thread handles 8 pieces (4 pixels of one frame): 16-byte chunk c = its index % 8, frames m = index / 8 + 16 q.
        // All raw loads of the NEXT group are issued right after the current stage is handed over and are consumed
        // together one group later (a single register set, in flight while the other three pairs work).
        const int ptid = tid - kPTEpiWarps * 32;
        constexpr int kGroup = kPTProducers / kPTTiles;            // producer threads per frame tile (128)
        constexpr int NQ = 8 * 128 / kGroup;                       // pieces per thread and stage (8)
        constexpr int kFStep = kGroup / 8;                         // frames between the pieces of a thread (16)
        const int ft = ptid / kGroup, idx = ptid % kGroup;
        const int c = idx & 7, m0 = idx >> 3;
        if (ft < nft) {
            // frame m = m0 + kFStep q: row m & 7 = m0 & 7 of 8-row atom (m0 >> 3) + (kFStep / 8) q
            const int srow = (m0 >> 3) * 1024 + (m0 & 7) * 128 + ((c ^ (m0 & 7)) << 4);
            constexpr int kQBytes = (kFStep / 8) * 1024;
            // frames past the end of the movie repeat the last frame (their results are never stored)
            const int64_t fa = f0 + 128 * ft + m0;
            const int qmax = fa < t ? (int)min((int64_t)(NQ - 1), (t - 1 - fa) / kFStep) : 0;
            const T* const fptr = movie + min(fa, t - 1) * d + it.c0 + 4 * c;
            const int64_t qstep = fa < t ? kFStep * d : 0;
            Px4<T> pre[NQ];
            float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = make_float4(1.f, 1.f, 1.f, 1.f);
            const bool ok_last = 32 * (it.nkc - 1) + 4 * c < 8 * it.w8;   // only the last chunk of a row can be partial
            auto issue_loads = [&](int64_t goff, bool ok) {
                if (!ok || (ablate & 2)) return;
                if (mean) mu = __ldg(reinterpret_cast<const float4*>(mean + goff + it.c0 + 4 * c));
                if (inv_std) is = __ldg(reinterpret_cast<const float4*>(inv_std + goff + it.c0 + 4 * c));
                const T* p = fptr + goff;
#pragma unroll
                for (int q = 0; q < NQ; ++q) pre[q].load(p + min(q, qmax) * qstep);
            };
            int kc = 0;
            int64_t goff = (int64_t)it.row0 * d2;                 // pixel offset of the group: row * d2 + 32 kc
            issue_loads(goff, it.nkc > 1 || ok_last);
            const int st = ft;                                     // stage = frame tile (phase = group: no parity aliasing)
            const uint32_t my_full = pt_smem_u32(&bar_afull[st]), my_empty = pt_smem_u32(&bar_aempty[st]);
            const uint32_t a_tf = sbase + st * kPTAStage + srow, a_bf = a_tf + kPTATile;
            for (int g = 0; g < n_groups; ++g) {
                const bool ok = kc < it.nkc - 1 || ok_last;
                if (g >= 1) pt_mbar_wait(my_empty, (g - 1) & 1);
                if (ok && !(ablate & 4)) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float4 x = pre[q].get();
                        const float y0 = (x.x - mu.x) * is.x, y1 = (x.y - mu.y) * is.y, y2 = (x.z - mu.z) * is.z, y3 = (x.w - mu.w) * is.w;
                        const float h0 = __uint_as_float(__float_as_uint(y0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(y1) & 0xFFFFE000u);
                        const float h2 = __uint_as_float(__float_as_uint(y2) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(y3) & 0xFFFFE000u);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a_tf + q * kQBytes), "f"(h0), "f"(h1), "f"(h2), "f"(h3)
                                     : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a_bf + q * kQBytes), "r"(pt_pack_bf16(h0, y0 - h0)),
                                     "r"(pt_pack_bf16(h1, y1 - h1)), "r"(pt_pack_bf16(h2, y2 - h2)), "r"(pt_pack_bf16(h3, y3 - h3))
                                     : "memory");
                    }
                }
                if (!(ablate & 16)) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                pt_mbar_arrive(my_full);
                if (++kc == it.nkc) {
                    kc = 0;
                    goff += d2 - 32 * (it.nkc - 1);
                } else {
                    goff += 32;
                }
                if (g + 1 < n_groups) issue_loads(goff, kc < it.nkc - 1 || ok_last);
            }
        }
    } else if (warp == kPTEpiWarps + kPTProdWarps) {
        // ================================ MMA issuer ================================
        // The whole warp runs the (warp-uniform) control flow so that addresses and descriptors stay in uniform
        // registers; one elected lane issues the tcgen05 instructions of a stage and its commits.
        // D f32, A / B K-major, N = 128, M = 128
        constexpr uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kPTN >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPTN >> 3) << 17) | ((128u >> 4) << 24);
        // K-major SWIZZLE_128B descriptor: low word = start address >> 4, high word = SBO 1024 >> 4 | version | layout type
        constexpr uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const PTEvent* ev = events + it.ev0;
        int e = 0, k = 0;
        // barrier addresses once (arrays of 8-byte barriers: + 8 per stage)
        const uint32_t afull0 = pt_smem_u32(&bar_afull[0]), aempty0 = pt_smem_u32(&bar_aempty[0]);
        const uint32_t bfull0 = pt_smem_u32(&bar_bfull[0]), bempty0 = pt_smem_u32(&bar_bempty[0]);
        const uint32_t accfull = pt_smem_u32(&bar_accfull), accfree = pt_smem_u32(&bar_accfree);
        const bool leader = pt_elect_one();
        const uint32_t a_tf0 = sbase >> 4, b_tf0 = sb_base >> 4;
        pt_mbar_wait(accfree, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        int kc = 0, row = it.row0, bs = 0;
        uint32_t bpar = 0;
        int next_ev_row = it.n_ev > 0 ? ev[0].row : -1;
        for (int g = 0; g < n_groups; ++g) {
            const int nks = min(4, it.w8 - 4 * kc);
            pt_mbar_wait(bfull0 + 8 * bs, bpar);
            const uint32_t b_tf = b_tf0 + bs * (kPTBStage >> 4), b_bf = b_tf + (kPTBPart >> 4);
#pragma unroll
            for (int ft = 0; ft < kPTTiles; ++ft) {
                if (ft < nft) {
                    pt_mbar_wait(afull0 + 8 * ft, g & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    if (leader) {
                        const uint32_t a_tf = a_tf0 + ft * (kPTAStage >> 4), a_bf = a_tf + (kPTATile >> 4);
                        const uint32_t dcol = tmem_d + kPTN * ft;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            if (ks < nks && !(ablate & 1)) {
                                pt_mma_tf32(dcol, desc_hi | (a_tf + 2 * ks), desc_hi | (b_tf + 2 * ks), idesc_tf32);
                                pt_mma_bf16(dcol, desc_hi | (a_bf + 2 * ks), desc_hi | (b_bf + 2 * ks), idesc_bf16);
                            }
                        }
                        pt_commit(aempty0 + 8 * ft);
                        if (ft == nft - 1) pt_commit(bempty0 + 8 * bs);
                    }
                }
            }
            if (++bs == kPTBStages) {
                bs = 0;
                bpar ^= 1;
            }
            if (kc == it.nkc - 1 && row == next_ev_row) {
                // tasks end at this row: let the epilogue warps drain and clear their slots
                while (e < it.n_ev && ev[e].row == row) ++e;
                next_ev_row = e < it.n_ev ? ev[e].row : -1;
                if (leader) pt_commit(accfull);
                ++k;
                pt_mbar_wait(accfree, k & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            }
            if (++kc == it.nkc) {
                kc = 0;
                ++row;
                if (leader) s_rows_issued = row - it.row0;
            }
        }
    } else if (warp == kPTEpiWarps + kPTProdWarps + 1) {
        // ================================ B loader (one thread) ================================
        if (lane == 0) {
            const unsigned char* src = bimg + (int64_t)it.b_chunk0 * kPTBStage;
            const uint32_t bfull0 = pt_smem_u32(&bar_bfull[0]), bempty0 = pt_smem_u32(&bar_bempty[0]);
            for (int g = 0; g < n_groups; ++g) {
                const int bs = g % kPTBStages;
                if (g + kPTBPrefetch < n_groups && !(ablate & 8))
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src + (int64_t)(g + kPTBPrefetch) * kPTBStage),
                                 "r"((uint32_t)kPTBStage)
                                 : "memory");
                if (g >= kPTBStages) pt_mbar_wait(bempty0 + 8 * bs, ((g / kPTBStages) - 1) & 1);
                const uint32_t bar = bfull0 + 8 * bs;
                if (ablate & 8) {
                    pt_mbar_arrive(bar);
                    continue;
                }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)kPTBStage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                 sb_base + bs * kPTBStage),
                             "l"(src + (int64_t)g * kPTBStage), "r"((uint32_t)kPTBStage), "r"(bar)
                             : "memory");
            }
        }
    }
    else if (pf_rows > 0) {
        // ================================ L2 prefetcher ================================
        // asks L2 for the 128-byte lines of the strip row the producers will load pf_rows rows after the row the MMA
        // stream is at (one line per lane and instruction), so that the producers' loads are L2 hits
        const int nframes = (int)min((int64_t)(128 * kPTTiles), t - f0);
        const int nlines = (int)((it.w8 * 8 * sizeof(T) + 127) / 128);
        for (int rr = 1; rr < it.n_rows; ++rr) {
            while (s_rows_issued + pf_rows < rr) __nanosleep(100);
            const char* p = reinterpret_cast<const char*>(movie + (int64_t)(it.row0 + rr) * d2 + it.c0);
            for (int f = lane; f < nframes; f += 32)
                for (int l = 0; l < nlines; ++l)   // the lines of one frame back to back: one DRAM row activation
                    asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p + (f0 + f) * d * (int64_t)sizeof(T) + 128 * l));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == kPTEpiWarps + kPTProdWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(kCols));
}

// Coefficient images: one CTA per (item, row); chunk kc of the row is the 32 KB shared-memory image
//   part 0: [128 slot columns][32 pixels] float32, TF32-exact hi values, SWIZZLE_128B
//   part 1: same shape, per pixel the bf16 pair (bf16(lo) in bits [0,16), bf16(hi) in [16,32))
__global__ void __launch_bounds__(256)
pack_strips_tc_kernel(const PTItem* __restrict__ items, const int32_t* __restrict__ item_of_row, const int32_t* __restrict__ slot_ptr,
                      const int32_t* __restrict__ tasks, const float* __restrict__ uvals, int64_t bpix, const float* __restrict__ bg,
                      int64_t d, int64_t d2, unsigned char* __restrict__ bimg) {
    __shared__ int s_task[kPTN / kPTSlotCols];
    const int ii = item_of_row[2 * blockIdx.x], rr = item_of_row[2 * blockIdx.x + 1];
    const PTItem it = items[ii];
    const int row = it.row0 + rr;
    if (threadIdx.x < kPTN / kPTSlotCols) {
        int found = -1;
        const int a = slot_ptr[it.slot_ptr0 + threadIdx.x], b = slot_ptr[it.slot_ptr0 + threadIdx.x + 1];
        for (int i = a; i < b; ++i) {
            const int by = tasks[8 * i], h = tasks[8 * i + 2];
            if (row >= by && row < by + h) found = i;
        }
        s_task[threadIdx.x] = found;
    }
    __syncthreads();
    unsigned char* out = bimg + ((int64_t)it.b_chunk0 + (int64_t)rr * it.nkc) * kPTBStage;
    const int pieces = it.nkc * kPTN * 8;     // 16-byte pieces (4 pixels of one column) of one part
    for (int p = threadIdx.x; p < pieces; p += blockDim.x) {
        const int c = p & 7, n = (p >> 3) % kPTN, kc = p / (8 * kPTN);
        const int ti = s_task[n / kPTSlotCols], comp = n % kPTSlotCols;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (ti >= 0) {
            const int32_t* tk = tasks + 8 * ti;
            if (comp < tk[5]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int x = 32 * kc + 4 * c + j - tk[1];
                    if (x >= 0 && x < tk[3]) {
                        v[j] = tk[6] == 0 ? uvals[(int64_t)(tk[4] + comp) * bpix + (int64_t)(row - tk[0]) * tk[3] + x]
                                          : bg[(int64_t)(tk[4] + comp) * d + (int64_t)row * d2 + it.c0 + tk[1] + x];
                    }
                }
            }
        }
        float hi[4];
        uint32_t pr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
            pr[j] = pt_pack_bf16(v[j] - hi[j], hi[j]);
        }
        const int64_t off = (int64_t)kc * kPTBStage + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4);
        *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(out + off + kPTBPart) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
    }
}

}  // namespace pmd

extern "C" int pmd_pack_strips_tc(const int32_t* items, const int32_t* item_of_row, int64_t n_rows_total, const int32_t* slot_ptr,
                                  const int32_t* tasks, const float* uvals, int64_t bpix, const float* bg, int64_t d, int64_t d2,
                                  void* bimg, void* stream) {
    const char* fn = "pmd_pack_strips_tc";
    PMD_REQUIRE(items && item_of_row && slot_ptr && tasks && bimg, fn, "null pointer");
    PMD_REQUIRE(n_rows_total > 0 && bpix > 0 && d > 0 && d2 > 0, fn, "bad size");
    PMD_REQUIRE(((uintptr_t)bimg & 15) == 0, fn, "image must be 16-byte aligned");
    pmd::pack_strips_tc_kernel<<<(unsigned)n_rows_total, 256, 0, (cudaStream_t)stream>>>(
        (const pmd::PTItem*)items, item_of_row, slot_ptr, tasks, uvals, bpix, bg, d, d2, (unsigned char*)bimg);
    return pmd::check_launch(fn);
}

extern "C" int pmd_project_stream_tc(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* items,
                                     int64_t n_items, const int32_t* events, const void* bimg, const float* mean,
                                     const float* inv_std, float* z, int64_t ldz, float* zbg, int64_t ldzbg, int64_t bg_stride,
                                     void* stream) {
    const char* fn = "pmd_project_stream_tc";
    PMD_REQUIRE(movie && items && events && bimg && z && zbg, fn, "null pointer");
    PMD_REQUIRE(t > 0 && n_items > 0 && ldz >= t && ldzbg >= t && d2 > 0 && d >= d2, fn, "bad size");
    PMD_REQUIRE((d2 & 3) == 0 && (d & 3) == 0, fn, "row length must be a multiple of 4 pixels");
    PMD_REQUIRE(((uintptr_t)movie & 15) == 0 && ((uintptr_t)bimg & 15) == 0, fn, "movie and image must be 16-byte aligned");
    PMD_REQUIRE((!mean || ((uintptr_t)mean & 15) == 0) && (!inv_std || ((uintptr_t)inv_std & 15) == 0), fn,
                "mean / inv_std must be 16-byte aligned");
    const int64_t ftiles = (t + 128 * pmd::kPTTiles - 1) / (128 * pmd::kPTTiles);
    PMD_REQUIRE(n_items <= 65535, fn, "too many strip items");
    cudaStream_t st = (cudaStream_t)stream;
    // Release builds take no switches from the environment.  Development builds (-DPMD_TUNE): PMD_TC_ABLATE is a profiling
    // aid (results are wrong when set: bit 0 no MMAs, 1 no movie loads, 2 no operand stores, 3 no coefficient copies, 4 no
    // proxy fence); PMD_TC_PREFETCH_ROWS = rows of the movie the prefetch warp asks L2 for ahead of the MMA stream (0 = off)
    int ablate = 0, pf_rows = 1;   // one row ahead: the strip row of a frame reaches DRAM as one burst (-4 % at C2)
#ifdef PMD_TUNE
    if (const char* e = getenv("PMD_TC_ABLATE")) ablate = atoi(e);
    if (const char* e = getenv("PMD_TC_PREFETCH_ROWS")) pf_rows = atoi(e);
#endif
    PMD_DISPATCH_DTYPE(dtype, fn, {
        auto k = pmd::project_tc_kernel<scalar_t>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pmd::kPTSmem);
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
        k<<<dim3((unsigned)ftiles, (unsigned)n_items), pmd::kPTThreads, pmd::kPTSmem, st>>>(
            (const scalar_t*)movie, t, d2, d, (const pmd::PTItem*)items, (const pmd::PTEvent*)events, (const unsigned char*)bimg,
            mean, inv_std, z, ldz, zbg, ldzbg, bg_stride, ablate, pf_rows);
    });
    return pmd::check_launch(fn);
}
