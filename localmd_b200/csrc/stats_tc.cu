// K1 (v3) on the 5th-generation tensor cores: per-pixel mean + Welch high-band noise estimate in ONE streaming pass over
// the movie, with the segment spectra computed as a tensor-core contraction.
//
// Replaces pmd_loader.py:203-291 and preprocessing_utils.py:10-40 of the reference: per 1024-frame chunk and pixel,
// welch(trace, noverlap=128) (periodic Hann window, 256-sample segments, hop 128, constant detrend, one-sided density,
// mean over segments), then sqrt(mean(0.5 * Pxx[65..128])).
//   * The Hann-windowed DFT of a constant is non-zero only at bins 0 and +-1: the per-segment mean removal cannot change
//     bins >= 65; a per-pixel offset c0 (the chunk's first frame) is subtracted only for rounding.
//   * One radix-2 decimation-in-frequency step in the time domain: with w[t + 128] = 1 - w[t],
//         a[t] = w[t] x[t] + (1 - w[t]) x[t + 128],   b[t] = w[t] x[t] - (1 - w[t]) x[t + 128],   t = 0..127,
//     the even bins of the windowed 256-point DFT are a 128-point transform of a and the odd bins one of b.  Only bins
//     65..128 are wanted: two real (64 x 128) matrices (rows = (cos, -sin) per bin, the Nyquist row scaled by sqrt(1/2)),
//     i.e. per segment and pixel 2 x 64 x 128 multiply-adds -- and those run on the tensor cores:
//         D[128 pixels x 64] = A[128 pixels x 128 frames] * B[128 frames x 64]        for the even and for the odd half.
//   * float32 accuracy from TF32 / bf16 MMAs as in K7: A = hi + lo with hi exact in TF32; one kind::tf32 MMA (hi * hi)
//     plus one kind::f16 MMA of K = 16 per 8 frames whose K pairs carry bf16(hi_a) * bf16(lo_b) and bf16(lo_a) * bf16(hi_b)
//     (dropped terms < 2^-18; measured against scipy's welch: 2e-6 also for movies with strong slow signals, where the
//     rounding of the DFT matrix alone would leak 2e-4).
// A persistent CTA walks (chunk, 128-pixel strip) units, strips fastest, so the CTAs that run together read neighbouring
// strips of the same frames (whole image rows reach DRAM as bursts).  Roles (10 warps: at most 3 per scheduler, so a
// thread may hold 168 registers):
//   * one thread feeds a ring of raw [16 frames x 128 pixels] stages with 2-D TMA boxes (cp.async.bulk.tensor);
//   * two converter groups (4 warps each, thread = pixel = tensor-memory lane) take alternate stages: centre, window,
//     butterfly with the same stage of the previous hop block -- which the thread KEEPS IN REGISTERS (64 floats), so the
//     raw ring is purely streaming -- split, and write both operands straight into tensor memory (tcgen05.st);
//   * one warp issues the MMAs (A from tensor memory, B = the constant matrices in shared memory, 128 KB, loaded once);
//     accumulators are double buffered (2 x 128 columns), the first MMA of a segment overwrites;
//   * the converter groups also square-and-sum the finished segment (tcgen05.ld; even bins group 0, odd bins group 1)
//     one stage into the next hop block, while that one accumulates.
//   sigma^2 = 1/(64*96*nseg) * sum_seg ( sum_{k=65..127} |X[k]|^2 + 0.5 |X[128]|^2 )      (96 = sum w^2)
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace pmd {

constexpr int kSTPix = 128, kSTStage = 16, kSTHop = 128, kSTChunk = 1024;
constexpr int kSTConvWarps = 8, kSTMmaWarp = kSTConvWarps, kSTTmaWarp = kSTMmaWarp + 1;
constexpr int kSTThreads = (kSTTmaWarp + 1) * 32;
constexpr int kSTMaxRaw = 16, kSTAStages = 4;
constexpr int kSTBBytes = 131072, kSTTabBytes = kSTBBytes + 512;
constexpr uint32_t kSTAccCols = 256, kSTAStageCols = 64;
constexpr int kSTSmemBudget = 227 * 1024 - 1024;

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_wait(uint32_t bar, uint32_t parity) {
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
            printf("stats_tc stuck: block %d thread %d barrier smem 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
#else
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "ST_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra ST_DONE;\n\t"
        "bra ST_WAIT;\n\t"
        "ST_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ bool st_elect_one() {
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
__device__ __forceinline__ void st_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void st_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void st_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]; acc == 0 overwrites D
__device__ __forceinline__ void st_mma_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void st_mma_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ uint32_t st_pack_bf16(float lo_half, float hi_half) {   // lo_half -> bits [0,16), hi_half -> [16,32)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi_half), "f"(lo_half));
    return r;
}
__device__ __forceinline__ void st_sttm16(uint32_t addr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(
            addr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// packed float32 pairs (FADD2 / FMUL2 of sm_100): half the issue slots of the conversion arithmetic
__device__ __forceinline__ uint64_t st_pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void st_unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t st_add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t st_sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t st_mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// x -> (TF32-exact hi, bf16 pair (bf16(hi), bf16(x - hi))) for both halves of a packed pair
__device__ __forceinline__ void st_split2(uint64_t x2, uint32_t& hi0, uint32_t& hi1, uint32_t& pr0, uint32_t& pr1) {
    float x0, x1, l0, l1;
    st_unpack2(x2, x0, x1);
    hi0 = __float_as_uint(x0) & 0xFFFFE000u;
    hi1 = __float_as_uint(x1) & 0xFFFFE000u;
    st_unpack2(st_sub2(x2, st_pack2(__uint_as_float(hi0), __uint_as_float(hi1))), l0, l1);
    pr0 = st_pack_bf16(__uint_as_float(hi0), l0);
    pr1 = st_pack_bf16(__uint_as_float(hi1), l1);
}

struct STUnit {
    int chunk, x0, n, nst, nseg;
    int64_t f_begin;
};
__device__ __forceinline__ STUnit st_unit(int unit, int n_strips, int64_t t_local) {
    STUnit u;
    u.chunk = unit / n_strips;
    u.x0 = (unit - u.chunk * n_strips) * kSTPix;
    u.f_begin = (int64_t)u.chunk * kSTChunk;
    u.n = (int)min((int64_t)kSTChunk, t_local - u.f_begin);
    u.nst = (u.n + kSTStage - 1) / kSTStage;
    u.nseg = u.n >= 256 ? (u.n - kSTHop) / kSTHop : 0;
    return u;
}

template <typename T>
__global__ void __launch_bounds__(kSTThreads, 1)
stats_tc_kernel(const __grid_constant__ CUtensorMap tm_movie, const T* __restrict__ movie, int64_t t_local, int64_t d, double inv_total,
                const unsigned char* __restrict__ tab, float* __restrict__ mean_part, float* __restrict__ noise_part, int n_strips,
                int n_units, int n_raw) {
    extern __shared__ __align__(1024) unsigned char stsm[];
    __shared__ __align__(8) uint64_t bar_rfull[kSTMaxRaw], bar_rempty[kSTMaxRaw], bar_afull[kSTAStages], bar_aempty[kSTAStages],
        bar_accfull[2], bar_accfree[2], bar_c0full[2], bar_c0empty[2], bar_mfull, bar_mempty, bar_b;
    __shared__ uint32_t tmem_base_s;
    constexpr int kSlotBytes = kSTStage * kSTPix * (int)sizeof(T);
    const uint32_t sbase = (st_smem_u32(stsm) + 1023u) & ~1023u;
    unsigned char* gbase = stsm + (sbase - st_smem_u32(stsm));
    const uint32_t sb_base = sbase;                                   // B image (1024-byte aligned atoms)
    const float* wtab = reinterpret_cast<const float*>(gbase + kSTBBytes);          // w[0..127]
    const T* c0buf = reinterpret_cast<const T*>(gbase + kSTBBytes + 512);           // 2 x 128 first-frame values (1 KB each)
    double* msum_s = reinterpret_cast<double*>(gbase + kSTBBytes + 512 + 2048);     // [2 groups][128]
    const uint32_t sr_base = sbase + kSTBBytes + 512 + 2048 + 2048;                 // raw ring
    const T* ring = reinterpret_cast<const T*>(gbase + kSTBBytes + 512 + 2048 + 2048);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kSTMaxRaw; ++s) {
            st_mbar_init(st_smem_u32(&bar_rfull[s]), 1);
            st_mbar_init(st_smem_u32(&bar_rempty[s]), 128);
        }
        for (int s = 0; s < kSTAStages; ++s) {
            st_mbar_init(st_smem_u32(&bar_afull[s]), 128);
            st_mbar_init(st_smem_u32(&bar_aempty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            st_mbar_init(st_smem_u32(&bar_accfull[s]), 1);
            st_mbar_init(st_smem_u32(&bar_accfree[s]), 32 * kSTConvWarps);
            st_mbar_init(st_smem_u32(&bar_c0full[s]), 1);
            st_mbar_init(st_smem_u32(&bar_c0empty[s]), 32 * kSTConvWarps);
        }
        st_mbar_init(st_smem_u32(&bar_mfull), 16 * kSTConvWarps);
        st_mbar_init(st_smem_u32(&bar_mempty), 16 * kSTConvWarps);
        st_mbar_init(st_smem_u32(&bar_b), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    if (warp == kSTMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(st_smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
    }
    if (warp == kSTTmaWarp && lane == 0) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&tm_movie) : "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const uint32_t tmem = tmem_base_s;

    if (warp < kSTMmaWarp) {
        // ================================ converters (+ the read-out of finished segments) ================================
        const int g = warp >> 2;                                      // group: takes the stages j = g, g + 2, ... of a hop block
        const int m = 32 * (warp & 3) + lane;                         // pixel within the strip = tensor-memory lane
        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t ta_lane = lane_base + kSTAccCols;
        st_mbar_wait(st_smem_u32(&bar_b), 0);                         // window table resident
        int s_base = 0, q_base = 0, seg_base = 0, ucnt = 0;           // raw stages / MMA stages / segments before this unit
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++ucnt) {
            const STUnit u = st_unit(unit, n_strips, t_local);
            const bool valid = (int64_t)u.x0 + m < d;
            const int cb = ucnt & 1;
            st_mbar_wait(st_smem_u32(&bar_c0full[cb]), (ucnt >> 1) & 1);
            const float c0 = valid ? to_f32(c0buf[cb * (1024 / (int)sizeof(T)) + m]) : 0.f;
            st_mbar_arrive(st_smem_u32(&bar_c0empty[cb]));
            uint64_t msum2 = 0ull;                                    // packed partial sums of (x - c0)
            float pw = 0.f;
            int cnt = 0, seg_read = 0;
            uint64_t p[4][8];                                         // w * (x - c0) of this thread's stages of the previous hop block
            const uint64_t c02 = st_pack2(c0, c0);
            int slot, use;                                            // raw stage of this thread's next stage (advances by 2)
            {
                const int sg0 = s_base + g;
                use = sg0 / n_raw;
                slot = sg0 - use * n_raw;
            }
            // sum of squares of this group's half (even bins: group 0, odd bins: group 1) of a finished segment
            auto read_segment = [&](int sidx) {
                const int segc = seg_base + sidx, buf = segc & 1;
                st_mbar_wait(st_smem_u32(&bar_accfull[buf]), (segc >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                float p0 = 0.f, p1 = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                        : "r"(lane_base + 128u * buf + 64u * g + 32u * c));
                    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        p0 = fmaf(__uint_as_float(v[i]), __uint_as_float(v[i]), p0);
                        p1 = fmaf(__uint_as_float(v[i + 1]), __uint_as_float(v[i + 1]), p1);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                st_mbar_arrive(st_smem_u32(&bar_accfree[buf]));
                pw += p0 + p1;
            };
            const int nblk = (u.nst + 7) >> 3;
            for (int blk = 0; blk < nblk; ++blk) {
                const bool mma = blk >= 1 && blk <= u.nseg;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 2 * jj + g, st = blk * 8 + j;
                    if (st < u.nst) {
                        st_mbar_wait(st_smem_u32(&bar_rfull[slot]), use & 1);
                        const T* src = ring + (size_t)slot * (kSTStage * kSTPix) + m;
                        const int left = u.n - kSTStage * st;          // frames of this stage inside the chunk
                        uint64_t xv[8];
                        // (pixels beyond the field of view read the zero fill of the TMA box and have c0 = 0)
                        if (left >= 16) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                xv[i] = st_sub2(st_pack2(to_f32(src[2 * i * kSTPix]), to_f32(src[(2 * i + 1) * kSTPix])), c02);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                xv[i] = st_pack2(2 * i < left ? to_f32(src[2 * i * kSTPix]) - c0 : 0.f,
                                                 2 * i + 1 < left ? to_f32(src[(2 * i + 1) * kSTPix]) - c0 : 0.f);
                        }
                        st_mbar_arrive(st_smem_u32(&bar_rempty[slot]));   // values are in registers
                        slot += 2;
                        if (slot >= n_raw) {
                            slot -= n_raw;
                            ++use;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) msum2 = st_add2(msum2, xv[i]);
                        cnt += min(left, 16);
                        const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(wtab + 16 * j);
                        if (!mma) {
                            if (blk == 0) {
#pragma unroll
                                for (int i4 = 0; i4 < 4; ++i4) {
                                    const ulonglong2 w = w2[i4];
                                    p[jj][2 * i4] = st_mul2(w.x, xv[2 * i4]);
                                    p[jj][2 * i4 + 1] = st_mul2(w.y, xv[2 * i4 + 1]);
                                }
                            }
                        } else {
                            uint64_t bb[8];
                            uint32_t hi[16], pr[16];
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                const ulonglong2 w = w2[i4];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const int i = 2 * i4 + e;
                                    const uint64_t pn = st_mul2(e ? w.y : w.x, xv[i]), q = st_sub2(xv[i], pn);   // w x, (1 - w) x
                                    const uint64_t a = st_add2(p[jj][i], q);
                                    bb[i] = st_sub2(p[jj][i], q);
                                    p[jj][i] = pn;
                                    st_split2(a, hi[2 * i], hi[2 * i + 1], pr[2 * i], pr[2 * i + 1]);
                                }
                            }
                            const int qa = q_base + (blk - 1) * 8 + j, as = qa & (kSTAStages - 1), ause = qa >> 2;
                            if (ause >= 1) {
                                st_mbar_wait(st_smem_u32(&bar_aempty[as]), (ause - 1) & 1);
                                asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                            }
                            const uint32_t ta = ta_lane + kSTAStageCols * as;
                            st_sttm16(ta, hi);
                            st_sttm16(ta + 16, pr);
#pragma unroll
                            for (int i = 0; i < 8; ++i) st_split2(bb[i], hi[2 * i], hi[2 * i + 1], pr[2 * i], pr[2 * i + 1]);
                            st_sttm16(ta + 32, hi);
                            st_sttm16(ta + 48, pr);
                            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
                            asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
                            st_mbar_arrive(st_smem_u32(&bar_afull[as]));
                        }
                    }
                    // the segment that was completed by the previous hop block has had a stage's time to finish
                    if (jj == 0 && seg_read < u.nseg && seg_read <= blk - 2) read_segment(seg_read++);
                }
            }
            while (seg_read < u.nseg) read_segment(seg_read++);
            s_base += u.nst;
            q_base += 8 * u.nseg;
            seg_base += u.nseg;
            // group 1 hands its partial sums to group 0, which writes the unit's results
            float msum_lo, msum_hi;
            st_unpack2(msum2, msum_lo, msum_hi);
            const float msum = msum_lo + msum_hi;
            if (g == 1) {
                if (ucnt >= 1) st_mbar_wait(st_smem_u32(&bar_mempty), (ucnt - 1) & 1);
                msum_s[m] = (double)c0 * (double)cnt + (double)msum;
                msum_s[kSTPix + m] = (double)pw;
                st_mbar_arrive(st_smem_u32(&bar_mfull));
            } else {
                st_mbar_wait(st_smem_u32(&bar_mfull), ucnt & 1);
                const double total = msum_s[m] + (double)c0 * (double)cnt + (double)msum;
                const float pwt = pw + (float)msum_s[kSTPix + m];
                st_mbar_arrive(st_smem_u32(&bar_mempty));
                const int64_t px = (int64_t)u.x0 + m;
                if (px < d) {
                    const int64_t o = (int64_t)u.chunk * d + px;
                    mean_part[o] = (float)(total * inv_total);
                    noise_part[o] = u.nseg > 0 ? sqrtf(pwt / (64.f * 96.f * (float)u.nseg)) : 0.f;
                }
            }
        }
    } else if (warp == kSTMmaWarp) {
        // ================================ MMA issuer ================================
        // D f32, A K-major from tensor memory, B K-major SWIZZLE_128B from shared memory, M = 128, N = 64
        constexpr uint32_t idesc_tf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint32_t idesc_bf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        constexpr uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
        const bool leader = st_elect_one();
        st_mbar_wait(st_smem_u32(&bar_b), 0);
        int segc = 0, qa = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const STUnit u = st_unit(unit, n_strips, t_local);
            for (int s = 0; s < u.nseg; ++s, ++segc) {
                const int buf = segc & 1, use = segc >> 1;
                if (use >= 1) {
                    st_mbar_wait(st_smem_u32(&bar_accfree[buf]), (use - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                }
                const uint32_t d_even = tmem + 128u * buf, d_odd = d_even + 64u;
#pragma unroll 1
                for (int j = 0; j < 8; ++j, ++qa) {
                    const int as = qa & (kSTAStages - 1);
                    st_mbar_wait(st_smem_u32(&bar_afull[as]), (qa >> 2) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
                    if (leader) {
                        const uint32_t a0 = tmem + kSTAccCols + kSTAStageCols * as;
                        // B: K atom j >> 1 (32 frames), 32-byte K step 2 (j & 1) + ks inside the 128-byte swizzle row
                        const uint32_t b0 = (sb_base + (uint32_t)(j >> 1) * 8192u + (uint32_t)(j & 1) * 64u) >> 4;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t acc = (j | ks) != 0;
                            st_mma_tf32(d_even, a0 + 8 * ks, desc_hi | (uint64_t)(b0 + 2 * ks), idesc_tf32, acc);
                            st_mma_bf16(d_even, a0 + 16 + 8 * ks, desc_hi | (uint64_t)(b0 + (65536u >> 4) + 2 * ks), idesc_bf16);
                            st_mma_tf32(d_odd, a0 + 32 + 8 * ks, desc_hi | (uint64_t)(b0 + (32768u >> 4) + 2 * ks), idesc_tf32, acc);
                            st_mma_bf16(d_odd, a0 + 48 + 8 * ks, desc_hi | (uint64_t)(b0 + (98304u >> 4) + 2 * ks), idesc_bf16);
                        }
                        st_commit(st_smem_u32(&bar_aempty[as]));
                        if (j == 7) st_commit(st_smem_u32(&bar_accfull[buf]));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================================ TMA producer (one thread) ================================
        if (lane == 0) {
            const uint32_t bb = st_smem_u32(&bar_b);
            st_expect_tx(bb, (uint32_t)kSTTabBytes);
            for (int c = 0; c < 4; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sb_base + c * 32768),
                             "l"(tab + c * 32768), "r"(32768u), "r"(bb)
                             : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(sb_base + kSTBBytes),
                         "l"(tab + kSTBBytes), "r"(512u), "r"(bb)
                         : "memory");
            uint64_t pol_stream;                                      // the movie is read once
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol_stream));
            int slot = 0, ucnt = 0;
            uint32_t use = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++ucnt) {
                const STUnit u = st_unit(unit, n_strips, t_local);
                const int cb = ucnt & 1;
                if (ucnt >= 2) st_mbar_wait(st_smem_u32(&bar_c0empty[cb]), ((ucnt >> 1) - 1) & 1);
                const uint32_t c0bytes = (uint32_t)(min((int64_t)kSTPix, d - u.x0) * (int64_t)sizeof(T));
                const uint32_t cbar = st_smem_u32(&bar_c0full[cb]);
                st_expect_tx(cbar, c0bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                 sb_base + kSTBBytes + 512 + cb * 1024),
                             "l"(movie + u.f_begin * d + u.x0), "r"(c0bytes), "r"(cbar)
                             : "memory");
                for (int st = 0; st < u.nst; ++st) {
                    if (use >= 1) st_mbar_wait(st_smem_u32(&bar_rempty[slot]), (use - 1) & 1);
                    const uint32_t bar = st_smem_u32(&bar_rfull[slot]);
                    st_expect_tx(bar, (uint32_t)kSlotBytes);
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(
                            sr_base + slot * kSlotBytes),
                        "l"(&tm_movie), "r"(u.x0), "r"((int)(u.f_begin + kSTStage * st)), "r"(bar), "l"(pol_stream)
                        : "memory");
                    if (++slot == n_raw) {
                        slot = 0;
                        ++use;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
    __syncthreads();
    if (warp == kSTMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u));
}

typedef CUresult (*STEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static STEncodeFn st_encode_fn() {
    static STEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (STEncodeFn)p;
    }();
    return fn;
}

template <typename T>
static int launch_stats_tc(const void* movie, int64_t t_local, int64_t d, int64_t t_total, const void* tab, float* mean_part,
                           float* noise_part, cudaStream_t st, const char* fn) {
    STEncodeFn enc = st_encode_fn();
    if (!enc) {
        set_error(std::string(fn) + ": cuTensorMapEncodeTiled is not available");
        return -2;
    }
    CUtensorMapDataType dt = sizeof(T) == 8   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                             : sizeof(T) == 4 ? (std::is_integral<T>::value ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32)
                             : sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                              : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)t_local};
    cuuint64_t gstr[1] = {(cuuint64_t)d * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)kSTPix, (cuuint32_t)kSTStage};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, dt, 2, const_cast<void*>(movie), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error(std::string(fn) + ": cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        return -3;
    }
    constexpr int kSlotBytes = kSTStage * kSTPix * (int)sizeof(T);
    constexpr int kFixed = kSTBBytes + 512 + 2048 + 2048;
    int n_raw = (kSTSmemBudget - kFixed) / kSlotBytes;
    if (n_raw > kSTMaxRaw) n_raw = kSTMaxRaw;
    if (n_raw < 3) return fail_arg(fn, "shared memory too small for this element type");
    const int smem = kFixed + n_raw * kSlotBytes + 1024;
    auto k = stats_tc_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
        set_error(std::string(fn) + ": " + cudaGetErrorString(e));
        return (int)e;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_strips = (d + kSTPix - 1) / kSTPix, n_chunks = (t_local + kSTChunk - 1) / kSTChunk;
    const int64_t n_units = n_strips * n_chunks;
    if (n_units > 0x7FFFFFFF) return fail_arg(fn, "too many (chunk, strip) units");
    const int grid = (int)(n_units < sms ? n_units : sms);
    k<<<grid, kSTThreads, smem, st>>>(tm, (const T*)movie, t_local, d, 1.0 / (double)t_total, (const unsigned char*)tab, mean_part,
                                      noise_part, (int)n_strips, (int)n_units, n_raw);
    return check_launch(fn);
}

}  // namespace pmd

extern "C" int pmd_stats_pass_tc(const void* movie, int dtype, int64_t t_local, int64_t d, int64_t t_total, const void* tab,
                                 float* mean_part, float* noise_part, void* stream) {
    const char* fn = "pmd_stats_pass_tc";
    PMD_REQUIRE(movie && tab && mean_part && noise_part, fn, "null pointer");
    PMD_REQUIRE(t_local > 0 && d > 0 && t_total > 0, fn, "non-positive size");
    PMD_REQUIRE(d < (1ll << 31) && t_local < (1ll << 31), fn, "movie too large for 32-bit TMA coordinates");
    PMD_REQUIRE(((uintptr_t)movie & 15) == 0 && ((uintptr_t)tab & 15) == 0, fn, "movie and table must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        PMD_REQUIRE(((uint64_t)d * sizeof(scalar_t)) % 16 == 0, fn, "frame pitch must be a multiple of 16 bytes (TMA)");
        return pmd::launch_stats_tc<scalar_t>(movie, t_local, d, t_total, tab, mean_part, noise_part, st, fn);
    });
    return 0;
}
