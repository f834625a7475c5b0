// Block projection with the movie operand in TENSOR MEMORY (the K7 recipe applied to the block stage):
//     out[b][c][f] = sum_q w[b][q][c] * yT[pix(b, q)][f]            (decomposition.py:295-298, 318, 390-407)
// Per (block, 128-frame tile): D[128 frames x 64 comps] = A[128 frames x K pixels] * B[K pixels x 64 comps].
//   * A: raw [32 pixel rows x 128 frames] tiles of the pixel-major init movie arrive by 2-D TMA boxes (a box = the
//     frames of 4 (or 2) consecutive pixels of a block row).  Two converter groups (thread = frame = tensor-memory lane)
//     split every value into hi (exact in TF32) and a bf16 pair (bf16(hi), bf16(x - hi)) and write the operand straight
//     into tensor memory (tcgen05.st): nothing is converted through shared memory (the previous kernel staged and split
//     A in shared memory -- 1 copy + 1 load + 2 stores per 16 bytes -- and was bound by that traffic).
//   * B: the block's components, packed once per call by block_pack_w_kernel into the K-major SWIZZLE_128B shared-
//     memory image (TF32 hi part + bf16 pair part (bf16(lo), bf16(hi)), 16 KB per 32 pixels), one bulk copy per chunk,
//     shared by the (up to 3) frame tiles of the unit.
//   * per 8 pixels one kind::tf32 MMA (hi * hi) + ONE kind::f16 MMA of K = 16 (hi * lo + lo * hi); dropped terms < 2^-18.
//   * persistent CTAs walk (block, frame group) units; the accumulators (3 tiles x 64 columns) are double buffered, so
//     the four epilogue warps store a finished unit while the next one accumulates.
#include <cuda.h>

#include "common.cuh"

namespace pmd {

constexpr int kBTEpiWarps = 4, kBTConvGroups = 2, kBTConvWarps = 4 * kBTConvGroups;
constexpr int kBTMmaWarp = kBTEpiWarps + kBTConvWarps, kBTTmaWarp = kBTMmaWarp + 1, kBTBWarp = kBTMmaWarp + 2;
constexpr int kBTThreads = (kBTBWarp + 1) * 32;
constexpr int kBTTiles = 3, kBTN = 64;
constexpr int kBTBStages = 3, kBTRawStages = 9, kBTAStages = 4;
constexpr int kBTBStageBytes = 2 * kBTN * 128;           // 16 KB: 32 pixels x 64 comps, TF32 part + pair part
constexpr int kBTRawBytes = 32 * 128 * 4;                 // 16 KB: 32 pixel rows x 128 frames
constexpr uint32_t kBTAccCols = 2 * kBTTiles * kBTN;      // 384: two buffers of 3 x 64
constexpr uint32_t kBTACols = 32;                         // one A stage: 16 pixels (hi 16 + pair 16 columns)

__device__ __forceinline__ uint32_t bt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bt_mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef PMD_TC_DEBUG
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
            printf("block_project_ts stuck: block %d thread %d barrier smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
#else
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "BT_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra BT_DONE;\n\t"
        "bra BT_WAIT;\n\t"
        "BT_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ bool bt_elect_one() {
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
__device__ __forceinline__ void bt_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bt_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bt_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bt_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bt_mma_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bt_mma_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ uint32_t bt_pack_bf16(float lo_half, float hi_half) {   // lo_half -> bits [0,16), hi_half -> [16,32)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}
__device__ __forceinline__ void bt_sttm16(uint32_t addr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(
            addr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

struct BTUnit {
    int b, nft;
    int64_t f0;
};
// Unit order: groups of `grp` consecutive blocks (about three rows of the block grid); inside a group the frame group is
// the slow index and the block the fast one.  The CTAs that run together then work on neighbouring blocks of the SAME
// frames (the overlapping pixels of neighbouring blocks are served by L2), and a group's coefficient images (~30 MB) stay
// in L2 across its frame groups.
__device__ __forceinline__ BTUnit bt_unit(int unit, int n_fg, int64_t ldo, int grp, int nb) {
    const int per_grp = grp * n_fg;
    const int g = unit / per_grp, rem = unit - g * per_grp;
    const int nbg = min(grp, nb - g * grp);                  // blocks of this group (the last one may be smaller)
    const int fg = rem / nbg;
    BTUnit u;
    u.b = g * grp + (rem - fg * nbg);
    u.f0 = (int64_t)fg * (128 * kBTTiles);
    u.nft = (int)min((int64_t)kBTTiles, (ldo - u.f0 + 127) / 128);
    return u;
}

// movie rows: row = b * rows_per_batch + pixel id; the tensor map is over [rows][ld] floats
__global__ void __launch_bounds__(kBTThreads, 1)
block_project_ts_kernel(const __grid_constant__ CUtensorMap tm_movie, int64_t rows_per_batch, int d2, const int32_t* __restrict__ starts,
                        int bw, int bpix, int box_h, const unsigned char* __restrict__ bimg, int r, float* __restrict__ out, int64_t ldo,
                        int n_fg, int n_units, int grp, int nb) {
    extern __shared__ __align__(1024) unsigned char btsm[];
    __shared__ __align__(8) uint64_t bar_rfull[kBTRawStages], bar_rempty[kBTRawStages], bar_afull[kBTAStages], bar_aempty[kBTAStages],
        bar_bfull[kBTBStages], bar_bempty[kBTBStages], bar_accfull[2], bar_accfree[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (bt_smem_u32(btsm) + 1023u) & ~1023u;
    unsigned char* gbase = btsm + (sbase - bt_smem_u32(btsm));
    const uint32_t sb_base = sbase;                                          // B stages
    const uint32_t sr_base = sbase + kBTBStages * kBTBStageBytes;            // raw stages
    const float* ring = reinterpret_cast<const float*>(gbase + kBTBStages * kBTBStageBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = (bpix + 31) >> 5;                                        // 32-pixel chunks per block

    if (tid == 0) {
        for (int s = 0; s < kBTRawStages; ++s) {
            bt_mbar_init(bt_smem_u32(&bar_rfull[s]), 1);
            bt_mbar_init(bt_smem_u32(&bar_rempty[s]), 128);
        }
        for (int s = 0; s < kBTAStages; ++s) {
            bt_mbar_init(bt_smem_u32(&bar_afull[s]), 128);
            bt_mbar_init(bt_smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < kBTBStages; ++s) {
            bt_mbar_init(bt_smem_u32(&bar_bfull[s]), 1);
            bt_mbar_init(bt_smem_u32(&bar_bempty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            bt_mbar_init(bt_smem_u32(&bar_accfull[s]), 1);
            bt_mbar_init(bt_smem_u32(&bar_accfree[s]), 32 * kBTEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    // raw stages start as zeros: the rows of a stage that no box writes (past the block's last pixel) must hold finite
    // values (they meet zero coefficients)
    for (int i = tid; i < kBTRawStages * kBTRawBytes / 16; i += kBTThreads)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(sr_base + 16 * i), "r"(0u) : "memory");
    if (warp == kBTMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(bt_smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    if (warp == kBTTmaWarp && lane == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_movie) : "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem = tmem_base_s;

    if (warp < kBTEpiWarps) {
        // ================================ epilogue: store a finished unit ================================
        const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
        int ucnt = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++ucnt) {
            const BTUnit u = bt_unit(unit, n_fg, ldo, grp, nb);
            const int buf = ucnt & 1;
            bt_mbar_wait(bt_smem_u32(&bar_accfull[buf]), (ucnt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            for (int ft = 0; ft < u.nft; ++ft) {
                const int64_t f = u.f0 + 128 * ft + 32 * warp + lane;
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    if (16 * cq < r) {
                        uint32_t v[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                            : "r"(lane_base + (uint32_t)(kBTTiles * kBTN) * buf + kBTN * ft + 16 * cq));
                        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                        if (f < ldo) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int c = 16 * cq + i;
                                if (c < r) out[((int64_t)u.b * r + c) * ldo + f] = __uint_as_float(v[i]);
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
            bt_mbar_arrive(bt_smem_u32(&bar_accfree[buf]));
        }
    } else if (warp < kBTMmaWarp) {
        // ================================ converters ================================
        // group j takes the work items (chunk, frame tile) i = j, j + 2, ... of the whole run; thread = frame
        const int j = (warp - kBTEpiWarps) >> 2;
        const int m = 32 * (warp & 3) + lane;
        const uint32_t ta0 = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + kBTAccCols + kBTACols * 2 * j;
        const uint32_t afull = bt_smem_u32(&bar_afull[2 * j]), aempty = bt_smem_u32(&bar_aempty[2 * j]);
        int64_t i_base = 0;                                               // items before this unit
        int n = 0;                                                        // items this group has converted
        int rs = j;                                                       // raw stage of this group's next item (advances by 2)
        uint32_t ruse = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const BTUnit u = bt_unit(unit, n_fg, ldo, grp, nb);
            const int n_items = nch * u.nft;
            const int first = (int)((j - (i_base & 1) + 2) & 1);         // first item of this unit whose global index has parity j
            for (int il = first; il < n_items; il += kBTConvGroups, ++n) {
                bt_mbar_wait(bt_smem_u32(&bar_rfull[rs]), ruse & 1);
                const float* src = ring + (size_t)rs * (kBTRawBytes / 4) + m;
                uint32_t hi[32], pr[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const float x = src[q * 128];
                    hi[q] = __float_as_uint(x) & 0xFFFFE000u;
                    pr[q] = bt_pack_bf16(__uint_as_float(hi[q]), x - __uint_as_float(hi[q]));
                }
                bt_mbar_arrive(bt_smem_u32(&bar_rempty[rs]));              // values are in registers
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (n >= 1) {
                        bt_mbar_wait(aempty + 8 * h, (n - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    }
                    uint32_t a[16], p[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        a[q] = hi[16 * h + q];
                        p[q] = pr[16 * h + q];
                    }
                    bt_sttm16(ta0 + kBTACols * h, a);
                    bt_sttm16(ta0 + kBTACols * h + 16, p);
                    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                    bt_mbar_arrive(afull + 8 * h);
                }
                rs += kBTConvGroups;
                if (rs >= kBTRawStages) {
                    rs -= kBTRawStages;
                    ++ruse;
                }
            }
            i_base += n_items;
        }
    } else if (warp == kBTMmaWarp) {
        // ================================ MMA issuer ================================
        // D f32, A K-major from tensor memory, B K-major SWIZZLE_128B from shared memory, M = 128, N = 64
        constexpr uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBTN >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBTN >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const bool leader = bt_elect_one();
        int64_t ig = 0;                                                   // global item counter
        int bs = 0, ucnt = 0;                                             // B stage of the next chunk, its use parity
        uint32_t bpar = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++ucnt) {
            const BTUnit u = bt_unit(unit, n_fg, ldo, grp, nb);
            const int buf = ucnt & 1;
            if (ucnt >= 2) {
                bt_mbar_wait(bt_smem_u32(&bar_accfree[buf]), ((ucnt >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            }
            const uint32_t d0 = tmem + (uint32_t)(kBTTiles * kBTN) * buf;
            for (int kc = 0; kc < nch; ++kc) {
                bt_mbar_wait(bt_smem_u32(&bar_bfull[bs]), bpar);
                const uint32_t b_tf = (sb_base + bs * kBTBStageBytes) >> 4, b_bf = b_tf + ((kBTN * 128) >> 4);
                for (int ft = 0; ft < u.nft; ++ft, ++ig) {
                    const uint32_t dcol = d0 + kBTN * ft;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int64_t ah = 2 * ig + h;
                        const int as = (int)(ah & (kBTAStages - 1));
                        bt_mbar_wait(bt_smem_u32(&bar_afull[as]), (uint32_t)(ah >> 2) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                        if (leader) {
                            const uint32_t a_hi = tmem + kBTAccCols + kBTACols * as, a_pr = a_hi + 16;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                bt_mma_tf32(dcol, a_hi + 8 * ks, desc_hi | (uint64_t)(b_tf + 2 * (2 * h + ks)), idesc_tf32,
                                            (uint32_t)((kc | h | ks) != 0));
                                bt_mma_bf16(dcol, a_pr + 8 * ks, desc_hi | (uint64_t)(b_bf + 2 * (2 * h + ks)), idesc_bf16);
                            }
                            bt_commit(bt_smem_u32(&bar_aempty[as]));
                            if (h == 1 && ft == u.nft - 1) bt_commit(bt_smem_u32(&bar_bempty[bs]));
                            if (h == 1 && ft == u.nft - 1 && kc == nch - 1) bt_commit(bt_smem_u32(&bar_accfull[buf]));
                        }
                        __syncwarp();
                    }
                }
                if (++bs == kBTBStages) {
                    bs = 0;
                    bpar ^= 1;
                }
            }
        }
    } else if (warp == kBTTmaWarp) {
        // ================================ raw movie tiles ================================
        // lane l issues the l-th box (4 or 2 pixel rows) of every raw tile: the box rows of a chunk are computed once per
        // lane and reused by the unit's frame tiles (a single thread issuing 8 boxes per tile was the kernel's bottleneck)
        int rs = 0;
        uint32_t ruse = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const BTUnit u = bt_unit(unit, n_fg, ldo, grp, nb);
            const int i0 = starts[2 * u.b], j0 = starts[2 * u.b + 1];
            const int64_t row_base = (int64_t)u.b * rows_per_batch;
            for (int kc = 0; kc < nch; ++kc) {
                const int q_lo = 32 * kc, q_hi = min(bpix, q_lo + 32);
                const uint32_t bytes = (uint32_t)(q_hi - q_lo) * 512u;
                const int q = q_lo + lane * box_h;
                const bool active = q < q_hi;
                const int qi = q / bw, qj = q - qi * bw;
                const int row = (int)(row_base + (int64_t)(i0 + qi) * d2 + j0 + qj);
                for (int ft = 0; ft < u.nft; ++ft) {
                    const uint32_t bar = bt_smem_u32(&bar_rfull[rs]);
                    if (lane == 0) {
                        if (ruse >= 1) bt_mbar_wait(bt_smem_u32(&bar_rempty[rs]), (ruse - 1) & 1);
                        bt_expect_tx(bar, bytes);
                    }
                    __syncwarp();
                    if (active)
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                                sr_base + rs * kBTRawBytes + (q - q_lo) * 512),
                            "l"(&tm_movie), "r"((int)(u.f0 + 128 * ft)), "r"(row), "r"(bar)
                            : "memory");
                    if (++rs == kBTRawStages) {
                        rs = 0;
                        ++ruse;
                    }
                }
            }
        }
    } else {
        // ================================ coefficient chunks (one thread) ================================
        if (lane == 0) {
            uint64_t pol_keep;                                            // shared by every frame group of the block
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol_keep));
            int bs = 0;
            uint32_t buse = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
                const BTUnit u = bt_unit(unit, n_fg, ldo, grp, nb);
                const unsigned char* src = bimg + (int64_t)u.b * nch * kBTBStageBytes;
                for (int kc = 0; kc < nch; ++kc) {
                    if (buse >= 1) bt_mbar_wait(bt_smem_u32(&bar_bempty[bs]), (buse - 1) & 1);
                    const uint32_t bar = bt_smem_u32(&bar_bfull[bs]);
                    bt_expect_tx(bar, (uint32_t)kBTBStageBytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                                     sb_base + bs * kBTBStageBytes),
                                 "l"(src + (int64_t)kc * kBTBStageBytes), "r"((uint32_t)kBTBStageBytes), "r"(bar), "l"(pol_keep)
                                 : "memory");
                    if (++bs == kBTBStages) {
                        bs = 0;
                        ++buse;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == kBTMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}

// Coefficient images: per block ceil(bpix / 32) chunks of 16 KB:
//   part 0: [64 comps][32 pixels] float32, TF32-exact hi values, K-major SWIZZLE_128B;  part 1: same shape, per element the
//   bf16 pair (bf16(lo) in bits [0,16), bf16(hi) in [16,32)).  One CTA per (chunk, block).
__global__ void __launch_bounds__(256)
block_pack_w_kernel(const float* __restrict__ w, int bpix, int rp, int r, unsigned char* __restrict__ bimg) {
    const int kc = blockIdx.x, nch = gridDim.x;
    const int64_t b = blockIdx.y;
    unsigned char* outp = bimg + ((int64_t)b * nch + kc) * kBTBStageBytes;
    const float* wb = w + b * (int64_t)bpix * rp;
    for (int p = threadIdx.x; p < kBTN * 8; p += blockDim.x) {            // 16-byte pieces (4 pixels of one component)
        const int c = p & 7, n = p >> 3;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = 32 * kc + 4 * c + j;
                if (q < bpix) v[j] = wb[(int64_t)q * rp + n];
            }
        }
        float hi[4];
        uint32_t pr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
            pr[j] = bt_pack_bf16(v[j] - hi[j], hi[j]);
        }
        const int off = (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4);
        *reinterpret_cast<float4*>(outp + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(outp + off + kBTN * 128) = make_uint4(pr[0], pr[1], pr[2], pr[3]);
    }
}

typedef CUresult (*BTEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static BTEncodeFn bt_encode_fn() {
    static BTEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (BTEncodeFn)p;
    }();
    return fn;
}

}  // namespace pmd

extern "C" int64_t pmd_block_project_ts_workspace_bytes(int64_t nb, int64_t bh, int64_t bw) {
    const int64_t nch = (bh * bw + 31) / 32;
    return nb * nch * (int64_t)pmd::kBTBStageBytes;
}

extern "C" int pmd_block_project_ts(const float* movie_t, int64_t movie_batch_stride, int64_t n_rows, int64_t ld, int64_t d2,
                                    const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w, int64_t r, int64_t rp,
                                    void* workspace, float* out, int64_t ldo, void* stream) {
    const char* fn = "pmd_block_project_ts";
    PMD_REQUIRE(movie_t && starts && w && out && workspace, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && ldo > 0 && ldo <= ld && nb > 0 && nb <= 65535 && r > 0 && rp >= r && rp <= 64, fn,
                "bad size (ld multiple of 4, ldo <= ld, r <= rp <= 64)");
    PMD_REQUIRE(bw % 2 == 0 && bh > 0, fn, "block width must be even (2-D TMA boxes of 2 or 4 pixel rows)");
    PMD_REQUIRE(movie_batch_stride % ld == 0 && n_rows > 0 && n_rows < (1ll << 31) && ld < (1ll << 31), fn,
                "batch stride must be a whole number of pixel rows; 32-bit TMA coordinates");
    PMD_REQUIRE(((uintptr_t)movie_t % 16) == 0 && ((uintptr_t)workspace % 16) == 0, fn, "operands must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int bpix = (int)(bh * bw), nch = (bpix + 31) / 32;
    pmd::block_pack_w_kernel<<<dim3((unsigned)nch, (unsigned)nb), 256, 0, st>>>(w, bpix, (int)rp, (int)r, (unsigned char*)workspace);
    int rc = pmd::check_launch(fn);
    if (rc) return rc;
    pmd::BTEncodeFn enc = pmd::bt_encode_fn();
    if (!enc) {
        pmd::set_error(std::string(fn) + ": cuTensorMapEncodeTiled is not available");
        return -2;
    }
    const int box_h = bw % 4 == 0 ? 4 : 2;
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n_rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_h};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(movie_t), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        pmd::set_error(std::string(fn) + ": cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")");
        return -3;
    }
    const int smem = pmd::kBTBStages * pmd::kBTBStageBytes + pmd::kBTRawStages * pmd::kBTRawBytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(pmd::block_project_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e));
        return (int)e;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_fg = (ldo + 128 * pmd::kBTTiles - 1) / (128 * pmd::kBTTiles), n_units = nb * n_fg;
    PMD_REQUIRE(n_units < (1ll << 31), fn, "too many (block, frame group) units");
    const int grid = (int)(n_units < sms ? n_units : sms);
    pmd::block_project_ts_kernel<<<grid, pmd::kBTThreads, smem, st>>>(tm, movie_batch_stride / ld, (int)d2, starts, (int)bw, bpix, box_h,
                                                                     (const unsigned char*)workspace, (int)r, out, ldo, (int)n_fg,
                                                                     (int)n_units, (int)(nb < sms ? nb : sms), (int)nb);
    return pmd::check_launch(fn);
}

// =========================================================================================================================
// Block spatial projection with the movie operand in tensor memory:
//     s[b][q][c] = sum_f yT[pix(b, q)][f] * v[b][c][f]                              (decomposition.py:304-306)
// Per (block, 128-pixel tile): D[128 pixels x 64 comps] = A[128 pixels x K frames] * B[K frames x 64 comps].
//   * A: thread = pixel = tensor-memory lane reads 32 consecutive frames of its own row of the pixel-major movie (one
//     128-byte line, 8 x ld.global.v4) straight into registers -- one item ahead of the conversion --, splits them into
//     TF32 hi + bf16 pair and writes the operand into tensor memory (tcgen05.st).  No shared memory on the movie's path.
//   * B: one warp converts the block's temporal rows (64 comps x 32 frames per chunk) into the K-major SWIZZLE_128B
//     image (TF32 hi part + bf16 pair part) in shared memory; the chunk is shared by the unit's (up to 4) pixel tiles.
//   * per 8 frames one kind::tf32 MMA + ONE kind::f16 MMA of K = 16 (as in block_project_ts_kernel); the accumulators of
//     the 4 tiles (256 columns) stay in tensor memory over the whole frame loop; persistent CTAs walk (block, 512-pixel
//     group) units.
// =========================================================================================================================
namespace pmd {

constexpr int kBSTiles = 4;
constexpr int kBSBStages = 6, kBSAStages = 8;
constexpr uint32_t kBSAccCols = kBSTiles * kBTN;          // 256
constexpr int kBSEpiWarps = 4, kBSConvWarps = 8;
constexpr int kBSMmaWarp = kBSEpiWarps + kBSConvWarps, kBSBWarp = kBSMmaWarp + 1, kBSBWarps = 2;
constexpr int kBSThreads = (kBSBWarp + kBSBWarps) * 32;

__device__ __forceinline__ float4 bs_ldg128(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kBSThreads, 1)
block_spatial_ts_kernel(const float* __restrict__ movT, int64_t mbs, int64_t ld, int d2, const int32_t* __restrict__ starts, int bw,
                        int bpix, const float* __restrict__ v, int64_t ldv, int r, int rp, float* __restrict__ s, int n_tg,
                        int n_units) {
    extern __shared__ __align__(1024) unsigned char bssm[];
    __shared__ __align__(8) uint64_t bar_afull[kBSAStages], bar_aempty[kBSAStages], bar_bfull[kBSBStages], bar_bempty[kBSBStages],
        bar_accfull, bar_accfree;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (bt_smem_u32(bssm) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = (int)((ldv + 31) >> 5);                               // 32-frame chunks

    if (tid == 0) {
        for (int st = 0; st < kBSAStages; ++st) {
            bt_mbar_init(bt_smem_u32(&bar_afull[st]), 128);
            bt_mbar_init(bt_smem_u32(&bar_aempty[st]), 1);
        }
        for (int st = 0; st < kBSBStages; ++st) {
            bt_mbar_init(bt_smem_u32(&bar_bfull[st]), 32);
            bt_mbar_init(bt_smem_u32(&bar_bempty[st]), 1);
        }
        bt_mbar_init(bt_smem_u32(&bar_accfull), 1);
        bt_mbar_init(bt_smem_u32(&bar_accfree), 32 * kBSEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (warp == kBSMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(bt_smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem = tmem_base_s;

    auto unit_of = [&](int unit, int& b, int& q0, int& nft) {
        b = unit / n_tg;
        q0 = (unit - b * n_tg) * (128 * kBSTiles);
        nft = min(kBSTiles, (bpix - q0 + 127) >> 7);
    };

    if (warp < kBSEpiWarps) {
        // ================================ epilogue ================================
        const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
        int ucnt = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++ucnt) {
            int b, q0, nft;
            unit_of(unit, b, q0, nft);
            bt_mbar_wait(bt_smem_u32(&bar_accfull), ucnt & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            for (int ft = 0; ft < nft; ++ft) {
                const int q = q0 + 128 * ft + 32 * warp + lane;
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    if (16 * cq < rp) {
                        uint32_t vv[16];
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                            : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3]), "=r"(vv[4]), "=r"(vv[5]), "=r"(vv[6]), "=r"(vv[7]),
                              "=r"(vv[8]), "=r"(vv[9]), "=r"(vv[10]), "=r"(vv[11]), "=r"(vv[12]), "=r"(vv[13]), "=r"(vv[14]), "=r"(vv[15])
                            : "r"(lane_base + kBTN * ft + 16 * cq));
                        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                        if (q < bpix) {
                            float* o = s + ((int64_t)b * bpix + q) * rp + 16 * cq;
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                if (16 * cq + i < rp)
                                    *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(vv[i]), __uint_as_float(vv[i + 1]),
                                                                                    __uint_as_float(vv[i + 2]), __uint_as_float(vv[i + 3]));
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
            bt_mbar_arrive(bt_smem_u32(&bar_accfree));
        }
    } else if (warp < kBSMmaWarp) {
        // ================================ converters ================================
        const int j = (warp - kBSEpiWarps) >> 2;
        const int m = 32 * (warp & 3) + lane;
        // A half-stages (16 frames each): item n of group j uses 4 (n & 1) + 2 j + h -- two items of buffering per group, so a
        // group converts one item ahead of the MMAs of its previous one (with one item per group the hand-over latency of
        // store -> MMA -> commit bounded the whole kernel)
        const uint32_t ta_lane = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + kBSAccCols;
        int64_t i_base = 0;
        int n = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            int b, q0, nft;
            unit_of(unit, b, q0, nft);
            const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
            const float* mv = movT + (int64_t)b * mbs;
            const float* rowp[kBSTiles];                                  // this thread's pixel row in every tile (nullptr: no pixel)
#pragma unroll
            for (int ft = 0; ft < kBSTiles; ++ft) {
                const int q = q0 + 128 * ft + m;
                const int qi = q / bw, qj = q - qi * bw;
                rowp[ft] = (ft < nft && q < bpix) ? mv + ((int64_t)(i0 + qi) * d2 + j0 + qj) * ld : nullptr;
            }
            const int n_items = nch * nft;
            const int first = (int)((j - (i_base & 1) + 2) & 1);
            float4 raw[8];
            auto load_item = [&](int il) {
                const int kc = il / nft, ft = il - kc * nft;
                const float* p = nullptr;
#pragma unroll
                for (int t4 = 0; t4 < kBSTiles; ++t4)
                    if (t4 == ft) p = rowp[t4];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int64_t f = 32 * (int64_t)kc + 4 * c;
                    raw[c] = (p != nullptr && f < ldv) ? bs_ldg128(p + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            if (first < n_items) load_item(first);
            for (int il = first; il < n_items; il += kBTConvGroups, ++n) {
                uint32_t hi[32], pr[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float x[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[4 * c + e] = __float_as_uint(x[e]) & 0xFFFFE000u;
                        pr[4 * c + e] = bt_pack_bf16(__uint_as_float(hi[4 * c + e]), x[e] - __uint_as_float(hi[4 * c + e]));
                    }
                }
                if (il + kBTConvGroups < n_items) load_item(il + kBTConvGroups);    // next item's line is in flight during the hand-over
                if (il + 3 * kBTConvGroups < n_items) {                              // and the line after the next two is asked into L2
                    const int il2 = il + 3 * kBTConvGroups, kc2 = il2 / nft, ft2 = il2 - kc2 * nft;
                    const float* p2 = nullptr;
#pragma unroll
                    for (int t4 = 0; t4 < kBSTiles; ++t4)
                        if (t4 == ft2) p2 = rowp[t4];
                    if (p2 != nullptr && 32 * (int64_t)kc2 < ldv) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p2 + 32 * (int64_t)kc2));
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int as = 4 * (n & 1) + 2 * j + h;
                    if (n >= 2) {
                        bt_mbar_wait(bt_smem_u32(&bar_aempty[as]), ((n >> 1) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    }
                    uint32_t a[16], p[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        a[q] = hi[16 * h + q];
                        p[q] = pr[16 * h + q];
                    }
                    bt_sttm16(ta_lane + kBTACols * as, a);
                    bt_sttm16(ta_lane + kBTACols * as + 16, p);
                    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                    bt_mbar_arrive(bt_smem_u32(&bar_afull[as]));
                }
            }
            i_base += n_items;
        }
    } else if (warp == kBSMmaWarp) {
        // ================================ MMA issuer ================================
        constexpr uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBTN >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBTN >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const bool leader = bt_elect_one();
        int64_t ig = 0;
        int bs = 0, ucnt = 0;
        uint32_t bpar = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++ucnt) {
            int b, q0, nft;
            unit_of(unit, b, q0, nft);
            if (ucnt >= 1) {
                bt_mbar_wait(bt_smem_u32(&bar_accfree), (ucnt - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            }
            for (int kc = 0; kc < nch; ++kc) {
                bt_mbar_wait(bt_smem_u32(&bar_bfull[bs]), bpar);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                const uint32_t b_tf = (sbase + bs * kBTBStageBytes) >> 4, b_bf = b_tf + ((kBTN * 128) >> 4);
                for (int ft = 0; ft < nft; ++ft, ++ig) {
                    const uint32_t dcol = tmem + kBTN * ft;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int as = 4 * (int)((ig >> 1) & 1) + 2 * (int)(ig & 1) + h;   // item ig = item ig >> 1 of group ig & 1
                        bt_mbar_wait(bt_smem_u32(&bar_afull[as]), (uint32_t)(ig >> 2) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                        if (leader) {
                            const uint32_t a_hi = tmem + kBSAccCols + kBTACols * as, a_pr = a_hi + 16;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                bt_mma_tf32(dcol, a_hi + 8 * ks, desc_hi | (uint64_t)(b_tf + 2 * (2 * h + ks)), idesc_tf32,
                                            (uint32_t)((kc | h | ks) != 0));
                                bt_mma_bf16(dcol, a_pr + 8 * ks, desc_hi | (uint64_t)(b_bf + 2 * (2 * h + ks)), idesc_bf16);
                            }
                            bt_commit(bt_smem_u32(&bar_aempty[as]));
                            if (h == 1 && ft == nft - 1) bt_commit(bt_smem_u32(&bar_bempty[bs]));
                            if (h == 1 && ft == nft - 1 && kc == nch - 1) bt_commit(bt_smem_u32(&bar_accfull));
                        }
                        __syncwarp();
                    }
                }
                if (++bs == kBSBStages) {
                    bs = 0;
                    bpar ^= 1;
                }
            }
        }
    } else {
        // ================================ temporal rows -> operand image (two warps, alternate chunks) ================================
        // piece id = lane + 32 i (i < 16): 16-byte frame chunk c = id % 8 of component n = id / 8.  A warp's next chunk is
        // requested as soon as the current one is converted, so its global-memory latency overlaps the MMAs of the chunk
        // the other warp delivers.
        const int wb = warp - kBSBWarp;
        int64_t gb_base = 0;                                              // chunks before this unit
        float4 x[16];
        const float* vb = nullptr;
        auto load_chunk = [&](int kc) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int id = lane + 32 * i, c = id & 7, nn = id >> 3;
                const int64_t f = 32 * (int64_t)kc + 4 * c;
                x[i] = (nn < r && f < ldv) ? bs_ldg128(vb + (int64_t)nn * ldv + f) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            int b, q0, nft;
            unit_of(unit, b, q0, nft);
            vb = v + (int64_t)b * r * ldv;
            const int first = (int)((wb - (gb_base & 1) + 2) & 1);        // first chunk of this unit whose global index has parity wb
            if (first < nch) load_chunk(first);
            for (int kc = first; kc < nch; kc += kBSBWarps) {
                const int64_t gb = gb_base + kc;
                const int bs = (int)(gb % kBSBStages);
                const uint32_t buse = (uint32_t)(gb / kBSBStages);
                if (buse >= 1) bt_mbar_wait(bt_smem_u32(&bar_bempty[bs]), (buse - 1) & 1);
                const uint32_t base = sbase + bs * kBTBStageBytes;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int id = lane + 32 * i, c = id & 7, nn = id >> 3;
                    const float xv[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
                    float hi[4];
                    uint32_t pr[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[e] = __uint_as_float(__float_as_uint(xv[e]) & 0xFFFFE000u);
                        pr[e] = bt_pack_bf16(xv[e] - hi[e], hi[e]);
                    }
                    const uint32_t off = base + (nn >> 3) * 1024 + (nn & 7) * 128 + ((c ^ (nn & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(off), "f"(hi[0]), "f"(hi[1]), "f"(hi[2]), "f"(hi[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(off + kBTN * 128), "r"(pr[0]), "r"(pr[1]), "r"(pr[2]), "r"(pr[3])
                                 : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                bt_mbar_arrive(bt_smem_u32(&bar_bfull[bs]));
                if (kc + kBSBWarps < nch) load_chunk(kc + kBSBWarps);
            }
            gb_base += nch;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == kBSMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}

}  // namespace pmd

extern "C" int pmd_block_spatial_ts(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2, const int32_t* starts,
                                    int64_t nb, int64_t bh, int64_t bw, const float* v, int64_t ldv, int64_t r, int64_t rp, float* s,
                                    void* stream) {
    const char* fn = "pmd_block_spatial_ts";
    PMD_REQUIRE(movie_t && starts && v && s, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && ldv > 0 && ldv % 4 == 0 && ldv <= ld && nb > 0 && r > 0 && rp >= r && rp % 4 == 0 && rp <= 64, fn,
                "bad size (ld, ldv multiples of 4, ldv <= ld, rp multiple of 4, r <= rp <= 64)");
    PMD_REQUIRE(((uintptr_t)movie_t % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)s % 16) == 0 && (movie_batch_stride % 4) == 0,
                fn, "operands must be 16-byte aligned");
    const int64_t bpix = bh * bw;
    const int64_t n_tg = (bpix + 128 * pmd::kBSTiles - 1) / (128 * pmd::kBSTiles), n_units = nb * n_tg;
    PMD_REQUIRE(n_units < (1ll << 31) && bpix < (1ll << 30), fn, "too many (block, pixel group) units");
    const int smem = pmd::kBSBStages * pmd::kBTBStageBytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(pmd::block_spatial_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e));
        return (int)e;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(n_units < sms ? n_units : sms);
    pmd::block_spatial_ts_kernel<<<grid, pmd::kBSThreads, smem, (cudaStream_t)stream>>>(
        movie_t, movie_batch_stride, ld, (int)d2, starts, (int)bw, (int)bpix, v, ldv, (int)r, (int)rp, s, (int)n_tg, (int)n_units);
    return pmd::check_launch(fn);
}
