// Block projection on the 5th-generation tensor cores (tcgen05, sm_100a), float32-accurate by 3xTF32:
//     out[b][c][f] = sum_q w[b][q][c] * yT[pix(b,q)][f]            (decomposition.py:295-298, 318)
// Per (block, 128-frame tile) this is D[128 frames x 64 comps] = A[128 x K] B[K x 64], K = block pixels.
//   * A (frames x pixels) is read from the pixel-major init movie, where the 128 frames of a pixel are
//     contiguous: in UMMA terms an MN-major operand.  All 128 threads copy it with 16-byte cp.async pieces
//     straight into the canonical MN-major layout of 32-bit operands (SWIZZLE_128B_BASE32B: 128-byte rows of 32
//     frames per pixel, atoms of 4 pixel rows, the 32-byte chunk index XORed with the row index), then split their
//     own pieces in place:
//         hi = x with the low 13 mantissa bits cleared (exactly representable in TF32),  lo = x - hi  (exact)
//   * B (pixels x comps) comes pre-split from the host as w_hi / w_lo ([nb][bpix][rp] float32, hi exactly
//     representable in TF32) and is copied into the same kind of layout (MN-major, comps contiguous).
//   * one elected thread issues, per 8-pixel K step, tcgen05.mma.kind::tf32 for A_hi B_hi, A_lo B_hi, A_hi B_lo
//     (the dropped A_lo B_lo term is 2^-22 relative) accumulating in float32 in TENSOR MEMORY (64 columns);
//     tcgen05.commit -> mbarrier releases the shared-memory stage (four 24 KB stages per CTA, two CTAs per SM:
//     the kernel needs ~100 KB of copies in flight per SM to cover the L2 latency);
//   * epilogue: tcgen05.ld (32 lanes x 32 bit x 16 columns per instruction), warp w owns TMEM lanes 32w..32w+31
//     = frames, and stores coalesced rows of out.
// No TMA descriptor is needed: the 16-byte cp.async granule is exactly the swizzle granule.
#include "common.cuh"

namespace pmd {

constexpr int kTCM = 128;                 // frames per accumulator tile (UMMA M)
constexpr int kTCTiles = 4;               // accumulator tiles per CTA: every B chunk is used for 4 x 128 frames
constexpr int kTCN = 64;                  // padded components (UMMA N)
constexpr int kTCK = 8;                   // pixels per shared-memory stage (one K step of 8)
constexpr int kTCStages = 3;              // stages in flight per CTA
constexpr int kTCThreads = 128;
constexpr int kTCGroupStride = (kTCK / 8) * 1024;   // bytes between 32-element M/N groups (descriptor LBO)
constexpr int kTCABytes = (kTCM / 32) * kTCGroupStride;   // 4 KB per tile and hi/lo part
constexpr int kTCBBytes = (kTCN / 32) * kTCGroupStride;   // 2 KB per hi/lo part
// stage = [tile][hi, lo] A tiles, then B_hi, B_lo = 36 KB
constexpr int kTCStageBytes = 2 * kTCTiles * kTCABytes + 2 * kTCBBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}

// MN-major operands of 32-bit elements use the SWIZZLE_128B_BASE32B layout: rows of 128 bytes (32 elements along
// M/N) per K index, swizzle atom = 4 K rows, the 32-byte chunk index of a row XORed with (k & 3).
// Descriptor of the 8 K rows (one tf32 MMA) starting at `addr`.
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);                    // start address      bits [0,14)
    d |= (uint64_t)(kTCGroupStride >> 4) << 16;                // leading byte offset: stride between M/N atoms [16,30)
    d |= (uint64_t)(512 >> 4) << 32;                           // stride byte offset: stride between K atoms    [32,46)
    d |= (uint64_t)1 << 46;                                    // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                                    // layout type SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity));
}

__device__ __forceinline__ bool tc_elect_one() {
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

// Warp roles: warps 0..3 = producers (cp.async + hi/lo split) and, at the end, the epilogue; warp 4 = TMEM allocation
// and the single MMA-issuing thread.  Stage hand-over is by mbarriers only (full: 128 producer arrivals; empty: one
// tcgen05.commit arrival), so the MMA thread runs decoupled from the producers.  A CTA owns kTCTiles = 4 accumulator
// tiles (512 frames, 256 TMEM columns): the B chunk of a stage is fetched once from L2 and used by all four.
__global__ void __launch_bounds__(kTCThreads + 32, 2)
block_project_tc_kernel(const float* __restrict__ movT, int64_t mbs, int64_t ld, int64_t d2, const int32_t* __restrict__ starts,
                        int bh, int bw, const float* __restrict__ w_hi, const float* __restrict__ w_lo, int r, int rp,
                        float* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(1024) unsigned char tcsm[];
    __shared__ __align__(8) uint64_t bar_full[kTCStages], bar_empty[kTCStages];
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(tcsm) + 1023u) & ~1023u;   // swizzle atoms want an aligned base
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.y;
    const int64_t f0 = (int64_t)blockIdx.x * (kTCM * kTCTiles);
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const float* mv = movT + b * mbs;
    const int bpix = bh * bw;
    const float* whb = w_hi + b * (int64_t)bpix * rp;
    const float* wlb = w_lo + b * (int64_t)bpix * rp;
    const int nch = (bpix + kTCK - 1) / kTCK;
    const int rp4 = rp / 4;
    constexpr uint32_t kCols = kTCN * kTCTiles;   // 256 TMEM columns

    if (tid == 0) {
        for (int s = 0; s < kTCStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar_full[s])), "r"(kTCThreads));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem_d = tmem_base_s;

    // instruction descriptor: D f32, A/B tf32, both MN-major, N = 64, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((kTCN >> 3) << 17) | ((kTCM >> 4) << 24);

    if (warp < 4) {
        // ================================ producers ================================
        // A pieces of this thread (8 per chunk): j -> accumulator tile j / 2, pixel row k = warp + 4 (j % 2) of the chunk,
        // 16-byte frame chunk c16 = lane.  B piece (one hi + one lo): pixel row k = tid / 16, component chunk n16 = tid % 16.
        // All addresses advance incrementally from chunk to chunk (chunks are issued in order, once each).
        constexpr int NPA = 2 * kTCTiles;
        auto piece_off = [](int k, int c16) {  // byte offset inside an operand tile
            return (c16 >> 3) * kTCGroupStride + k * 128 + (((((c16 & 7) >> 1) ^ (k & 3)) << 5) | ((c16 & 1) << 4));
        };
        const int64_t step_px = (int64_t)kTCK * ld, wrap_px = (int64_t)(d2 - bw) * ld;
        const float* pa[2];      // row pointers of the two pixel rows at frame f0 + 4 lane (tile offsets are added on use)
        int qja[2], qa[2];
        uint32_t doff[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = warp + 4 * h;
            const int qi = k / bw, qj = k - qi * bw;
            qa[h] = k;
            qja[h] = qj;
            pa[h] = mv + ((int64_t)(i0 + qi) * d2 + j0 + qj) * ld + f0 + 4 * lane;
            doff[h] = piece_off(k, lane);
        }
        bool f_ok[kTCTiles];
#pragma unroll
        for (int tl = 0; tl < kTCTiles; ++tl) f_ok[tl] = f0 + kTCM * tl + 4 * lane < ld;
        const int n16 = tid & 15;
        const bool n_ok = n16 < rp4;
        int qb = tid >> 4;
        const float* pbh = whb + (int64_t)qb * rp + 4 * n16;
        const float* pbl = wlb + (int64_t)qb * rp + 4 * n16;
        const uint32_t boff = piece_off(tid >> 4, n16);

        auto issue = [&](int st) {   // issues the NEXT chunk (internal cursor) into stage st
            const uint32_t a0 = sbase + st * kTCStageBytes, b_hi = a0 + 2 * kTCTiles * kTCABytes, b_lo = b_hi + kTCBBytes;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int tl = 0; tl < kTCTiles; ++tl) {
                    const uint32_t dst = a0 + tl * 2 * kTCABytes + doff[h];
                    if (qa[h] < bpix && f_ok[tl]) {
                        tc_cp_async16(dst, pa[h] + kTCM * tl);
                    } else {
                        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(dst), "f"(0.f));
                    }
                }
                qa[h] += kTCK;
                qja[h] += kTCK;
                pa[h] += step_px;
                while (qja[h] >= bw) {
                    qja[h] -= bw;
                    pa[h] += wrap_px;
                }
            }
            if (qb < bpix && n_ok) {
                tc_cp_async16(b_hi + boff, pbh);
                tc_cp_async16(b_lo + boff, pbl);
            } else {
                asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(b_hi + boff), "f"(0.f));
                asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(b_lo + boff), "f"(0.f));
            }
            qb += kTCK;
            pbh += (int64_t)kTCK * rp;
            pbl += (int64_t)kTCK * rp;
        };
        // split this thread's own A pieces of stage st: hi in place, lo to the tile next to it
        auto split = [&](int st) {
            const uint32_t a0 = sbase + st * kTCStageBytes;
#pragma unroll
            for (int j = 0; j < NPA; ++j) {
                const uint32_t a_hi = a0 + (j >> 1) * 2 * kTCABytes + doff[j & 1], a_lo = a_hi + kTCABytes;
                float x0, x1, x2, x3;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(a_hi));
                const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
                const float h2 = __uint_as_float(__float_as_uint(x2) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(x3) & 0xFFFFE000u);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a_hi), "f"(h0), "f"(h1), "f"(h2), "f"(h3));
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(a_lo), "f"(x0 - h0), "f"(x1 - h1), "f"(x2 - h2),
                             "f"(x3 - h3));
            }
        };

#pragma unroll
        for (int c = 0; c < kTCStages - 1; ++c) {
            if (c < nch) issue(c);
            asm volatile("cp.async.commit_group;\n" ::);
        }
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch % kTCStages;
            const int nx = ch + kTCStages - 1;                    // chunk to prefetch: it reuses the stage of chunk ch-1
            if (nx < nch) {
                if (ch >= 1) mbar_wait(smem_u32(&bar_empty[(ch - 1) % kTCStages]), ((ch - 1) / kTCStages) & 1);
                issue(nx % kTCStages);
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group %0;\n" ::"n"(kTCStages - 1));   // this thread's pieces of chunk ch have landed
            split(st);
            asm volatile("fence.proxy.async.shared::cta;\n" ::);  // generic-proxy writes -> visible to the tensor core
            mbar_arrive(smem_u32(&bar_full[st]));
        }
        // ================================ epilogue ================================
        // the commit of the last chunk covers every earlier MMA; TMEM lane = frame within the tile, column = component
        mbar_wait(smem_u32(&bar_empty[(nch - 1) % kTCStages]), ((nch - 1) / kTCStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
#pragma unroll 1
        for (int tl = 0; tl < kTCTiles; ++tl) {
            const int64_t f = f0 + kTCM * tl + 32 * warp + lane;
            if (f0 + kTCM * tl >= ldo) break;
#pragma unroll
            for (int cq = 0; cq < 4; ++cq) {
                if (16 * cq >= r) break;
                uint32_t v[16];
                const uint32_t taddr = tmem_d + ((uint32_t)(32 * warp) << 16) + kTCN * tl + 16 * cq;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
                if (f < ldo) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = 16 * cq + i;
                        if (c < r) out[(b * r + c) * ldo + f] = __uint_as_float(v[i]);
                    }
                }
            }
        }
    } else {
        // ================================ MMA issuer ================================
        // the whole warp runs the (uniform) loop so that addresses and descriptors stay in uniform registers; one
        // elected lane issues the MMAs and the commit of a stage
        // MN-major SWIZZLE_128B_BASE32B descriptor: low word = start address >> 4 | LBO << 16, high word = SBO | version | type
        constexpr uint64_t desc_hi = (uint64_t)((512u >> 4) | (1u << 14) | (1u << 29)) << 32;
        constexpr uint32_t lbo = (uint32_t)(kTCGroupStride >> 4) << 16;
        const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
        const bool leader = tc_elect_one();
        int st = 0;
        uint32_t par = 0;
        for (int ch = 0; ch < nch; ++ch) {
            mbar_wait(full0 + 8 * st, par);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            if (leader) {
                const uint32_t a0 = ((sbase + st * kTCStageBytes) >> 4) | lbo;
                const uint32_t b_hi = a0 + ((2 * kTCTiles * kTCABytes) >> 4), b_lo = b_hi + (kTCBBytes >> 4);
                const uint64_t dbh = desc_hi | b_hi, dbl = desc_hi | b_lo;
#pragma unroll
                for (int tl = 0; tl < kTCTiles; ++tl) {
                    const uint64_t dah = desc_hi | (a0 + ((tl * 2 * kTCABytes) >> 4));
                    const uint64_t dal = dah + (kTCABytes >> 4);
                    umma_tf32(tmem_d + kTCN * tl, dah, dbl, idesc, ch != 0);
                    umma_tf32(tmem_d + kTCN * tl, dal, dbh, idesc, 1u);
                    umma_tf32(tmem_d + kTCN * tl, dah, dbh, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(empty0 + 8 * st)
                             : "memory");
            }
            if (++st == kTCStages) {
                st = 0;
                par ^= 1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(kCols));
}

}  // namespace pmd

extern "C" int pmd_block_project_tc(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                                    const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w_hi,
                                    const float* w_lo, int64_t r, int64_t rp, float* out, int64_t ldo, void* stream) {
    const char* fn = "pmd_block_project_tc";
    PMD_REQUIRE(movie_t && starts && w_hi && w_lo && out, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && ldo > 0 && ldo <= ld && nb > 0 && nb <= 65535 && r > 0 && rp >= r && rp % 4 == 0 && rp <= 64,
                fn, "bad size (ld multiple of 4, ldo <= ld, rp multiple of 4, r <= rp <= 64)");
    PMD_REQUIRE(((uintptr_t)movie_t % 16) == 0 && ((uintptr_t)w_hi % 16) == 0 && ((uintptr_t)w_lo % 16) == 0 &&
                    (movie_batch_stride % 4) == 0,
                fn, "operands must be 16-byte aligned");
    const size_t smem = pmd::kTCStages * pmd::kTCStageBytes + 1024;   // + slack for the 1024-byte alignment of the swizzle atoms
    cudaError_t e = cudaFuncSetAttribute(pmd::block_project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    dim3 grid((unsigned)((ldo + pmd::kTCM * pmd::kTCTiles - 1) / (pmd::kTCM * pmd::kTCTiles)), (unsigned)nb);
    pmd::block_project_tc_kernel<<<grid, pmd::kTCThreads + 32, smem, (cudaStream_t)stream>>>(
        movie_t, movie_batch_stride, ld, d2, starts, (int)bh, (int)bw, w_hi, w_lo, (int)r, (int)rp, out, ldo);
    return pmd::check_launch(fn);
}

// =================================================================================================
// Block spatial projection on tcgen05 (3xTF32):  s[b][q][c] = sum_f yT[pix(b,q)][f] * v[b][c][f]
// (decomposition.py:304-306).  Per (block, 256-pixel tile) D[2 x 128 pixels][64 comps] = A[pixels x frames] B[frames x comps]:
// both operands are K-major here (frames are contiguous for a pixel of yT and for a component of v), the canonical
// SWIZZLE_128B layout: one 128-byte row (32 frames) per pixel / component, atoms of 8 rows, the 16-byte chunk index
// XORed with the row index.  Producers (8 warps) copy A and B with 16-byte cp.async and split BOTH into hi / lo in
// shared memory; one thread issues the MMAs (per 32-frame stage: 4 K steps x 2 pixel tiles x 3 products); the 128
// accumulator columns live in tensor memory for the whole K loop (all frames).
// =================================================================================================
namespace pmd {

constexpr int kSTTiles = 2;                         // 128-pixel accumulator tiles per CTA
constexpr int kSTK = 32;                            // frames per stage (one 128-byte swizzle row)
constexpr int kSTStages = 2;
constexpr int kSTProducers = 256;
constexpr int kSTABytes = 128 * 128;                // one A tile part (128 pixels x 32 frames)
constexpr int kSTBBytes = 64 * 128;                 // one B part (64 comps x 32 frames)
constexpr int kSTStageBytes = 2 * kSTTiles * kSTABytes + 2 * kSTBBytes;   // 80 KB

// K-major SWIZZLE_128B descriptor (rows of 128 bytes, 8-row atoms 1024 bytes apart)
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;                          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                                    // descriptor version
    d |= (uint64_t)2 << 61;                                    // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(kSTProducers + 32, 1)
block_spatial_tc_kernel(const float* __restrict__ movT, int64_t mbs, int64_t ld, int64_t d2, const int32_t* __restrict__ starts,
                        int bh, int bw, const float* __restrict__ v, int64_t ldv, int r, int rp, float* __restrict__ s) {
    extern __shared__ __align__(1024) unsigned char stsm[];
    __shared__ __align__(8) uint64_t bar_full[kSTStages], bar_empty[kSTStages];
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(stsm) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.y;
    const int q0 = blockIdx.x * (128 * kSTTiles);
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const float* mv = movT + b * mbs;
    const float* vb = v + b * (int64_t)r * ldv;
    const int bpix = bh * bw;
    const int nch = (int)((ldv + kSTK - 1) / kSTK);
    constexpr uint32_t kCols = 64 * kSTTiles;

    if (tid == 0) {
        for (int st = 0; st < kSTStages; ++st) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar_full[st])), "r"(kSTProducers));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar_empty[st])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem_d = tmem_base_s;
    // D f32, A/B tf32, both K-major, N = 64, M = 128
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

    if (warp < 8) {
        // ================================ producers ================================
        // A pieces: id = tid + 256 j (j < 8): 16-byte frame chunk c = id % 8, row = id / 8 (tile = row / 128, m = row % 128)
        // B pieces: id = tid + 256 j (j < 2): chunk c = id % 8, component n = id / 8
        const int c = tid & 7;
        const float* pa[8];
        uint32_t aoff[8];
        bool a_ok[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int row = (tid >> 3) + 32 * j;
            const int tile = row >> 7, m = row & 127;
            const int q = q0 + row;
            a_ok[j] = q < bpix;
            const int qq = a_ok[j] ? q : 0;
            const int qi = qq / bw, qj = qq - qi * bw;
            pa[j] = mv + ((int64_t)(i0 + qi) * d2 + j0 + qj) * ld + 4 * c;
            aoff[j] = tile * (2 * kSTABytes) + (m >> 3) * 1024 + (m & 7) * 128 + ((c ^ (m & 7)) << 4);
        }
        const float* pb[2];
        uint32_t boff[2];
        bool b_ok[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int n = (tid >> 3) + 32 * j;
            b_ok[j] = n < r;
            pb[j] = vb + (int64_t)(b_ok[j] ? n : 0) * ldv + 4 * c;
            boff[j] = 2 * kSTTiles * kSTABytes + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4);
        }
        int fcur = 0;   // first frame of the next chunk to issue
        auto issue = [&](int st) {
            const uint32_t base = sbase + st * kSTStageBytes;
            const bool f_ok = fcur + 4 * c < ldv;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (a_ok[j] && f_ok) tc_cp_async16(base + aoff[j], pa[j] + fcur);
                else asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(base + aoff[j]), "f"(0.f));
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (b_ok[j] && f_ok) tc_cp_async16(base + boff[j], pb[j] + fcur);
                else asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};\n" ::"r"(base + boff[j]), "f"(0.f));
            }
            fcur += kSTK;
        };
        auto split_piece = [&](uint32_t hi, uint32_t lo) {
            float x0, x1, x2, x3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(hi));
            const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
            const float h2 = __uint_as_float(__float_as_uint(x2) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(x3) & 0xFFFFE000u);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(hi), "f"(h0), "f"(h1), "f"(h2), "f"(h3));
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(lo), "f"(x0 - h0), "f"(x1 - h1), "f"(x2 - h2), "f"(x3 - h3));
        };
        auto split = [&](int st) {
            const uint32_t base = sbase + st * kSTStageBytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) split_piece(base + aoff[j], base + aoff[j] + kSTABytes);
#pragma unroll
            for (int j = 0; j < 2; ++j) split_piece(base + boff[j], base + boff[j] + kSTBBytes);
        };

        issue(0);
        asm volatile("cp.async.commit_group;\n" ::);
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch & 1;
            if (ch + 1 < nch) {
                if (ch >= 1) mbar_wait(smem_u32(&bar_empty[st ^ 1]), ((ch - 1) >> 1) & 1);
                issue(st ^ 1);
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 1;\n" ::);
            split(st);
            asm volatile("fence.proxy.async.shared::cta;\n" ::);
            mbar_arrive(smem_u32(&bar_full[st]));
        }
        // ================================ epilogue ================================
        mbar_wait(smem_u32(&bar_empty[(nch - 1) & 1]), ((nch - 1) >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        const int tile = warp >> 2, quarter = warp & 3;       // a warp reads the TMEM lanes 32 (warp % 4) .. +31
        const int q = q0 + 128 * tile + 32 * quarter + lane;
#pragma unroll
        for (int cq = 0; cq < 4; ++cq) {
            if (16 * cq >= rp) break;
            uint32_t vv[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * quarter) << 16) + 64 * tile + 16 * cq;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(vv[0]), "=r"(vv[1]), "=r"(vv[2]), "=r"(vv[3]), "=r"(vv[4]), "=r"(vv[5]), "=r"(vv[6]), "=r"(vv[7]), "=r"(vv[8]),
                  "=r"(vv[9]), "=r"(vv[10]), "=r"(vv[11]), "=r"(vv[12]), "=r"(vv[13]), "=r"(vv[14]), "=r"(vv[15])
                : "r"(taddr));
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
    } else {
        // ================================ MMA issuer ================================
        // whole warp, uniform control flow; one elected lane issues (see block_project_tc_kernel)
        constexpr uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;   // K-major SWIZZLE_128B
        const uint32_t full0 = smem_u32(&bar_full[0]), empty0 = smem_u32(&bar_empty[0]);
        const bool leader = tc_elect_one();
        for (int ch = 0; ch < nch; ++ch) {
            const int st = ch & 1;
            mbar_wait(full0 + 8 * st, (ch >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
            if (leader) {
                const uint32_t base = (sbase + st * kSTStageBytes) >> 4;
                const uint32_t b_hi = base + ((2 * kSTTiles * kSTABytes) >> 4), b_lo = b_hi + (kSTBBytes >> 4);
#pragma unroll
                for (int ks = 0; ks < kSTK / 8; ++ks) {
                    const uint64_t dbh = desc_hi | (b_hi + 2 * ks), dbl = desc_hi | (b_lo + 2 * ks);
#pragma unroll
                    for (int tl = 0; tl < kSTTiles; ++tl) {
                        const uint32_t a_hi = base + ((tl * 2 * kSTABytes) >> 4), a_lo = a_hi + (kSTABytes >> 4);
                        const uint64_t dah = desc_hi | (a_hi + 2 * ks), dal = desc_hi | (a_lo + 2 * ks);
                        umma_tf32(tmem_d + 64 * tl, dah, dbl, idesc, (ch | ks) != 0);
                        umma_tf32(tmem_d + 64 * tl, dal, dbh, idesc, 1u);
                        umma_tf32(tmem_d + 64 * tl, dah, dbh, idesc, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(empty0 + 8 * st)
                             : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(kCols));
}

}  // namespace pmd

extern "C" int pmd_block_spatial_tc(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                                    const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* v, int64_t ldv,
                                    int64_t r, int64_t rp, float* s, void* stream) {
    const char* fn = "pmd_block_spatial_tc";
    PMD_REQUIRE(movie_t && starts && v && s, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && ldv > 0 && ldv % 4 == 0 && ldv <= ld && nb > 0 && nb <= 65535 && r > 0 && rp >= r &&
                    rp % 4 == 0 && rp <= 64,
                fn, "bad size (ld, ldv multiples of 4, ldv <= ld, rp multiple of 4, r <= rp <= 64)");
    PMD_REQUIRE(((uintptr_t)movie_t % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)s % 16) == 0 &&
                    (movie_batch_stride % 4) == 0,
                fn, "operands must be 16-byte aligned");
    const size_t smem = pmd::kSTStages * pmd::kSTStageBytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(pmd::block_spatial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    const int64_t bpix = bh * bw;
    dim3 grid((unsigned)((bpix + 128 * pmd::kSTTiles - 1) / (128 * pmd::kSTTiles)), (unsigned)nb);
    pmd::block_spatial_tc_kernel<<<grid, pmd::kSTProducers + 32, smem, (cudaStream_t)stream>>>(
        movie_t, movie_batch_stride, ld, d2, starts, (int)bh, (int)bw, v, ldv, (int)r, (int)rp, s);
    return pmd::check_launch(fn);
}
