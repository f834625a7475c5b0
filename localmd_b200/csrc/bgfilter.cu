// Background removal on the pixel-major init movie (pmd_loader.py:374-389: standardize_and_filter):
//     vbg[k][f] = sum_p bg[k][p] * yt[p][f]          (K <= 16 dense background components)
//     yt[p][f] -= sum_k bg[k][p] * vbg[k][f]
// Both are skinny contractions (K = 15 at the defaults) bound by one pass over yt (5.2 GB at C2); the library GEMMs
// picked for these shapes take 2-3x the streaming time.  A thread owns 4 consecutive frames and keeps its K x 4
// accumulators (pass 1) or its K x 4 slice of vbg (pass 2) in registers while it walks a range of pixels; the K
// coefficients of a pixel are broadcast from shared memory.  Pass 1 writes one partial per pixel range (summed in a
// fixed order by the caller: deterministic, no atomics).
#include "common.cuh"

namespace pmd {

constexpr int kBGK = 16;
constexpr int kBGThreads = 256;
constexpr int kBGPix = 256;   // pixels staged per shared-memory refill
constexpr int kBGBatch = 4;   // pixel rows per pipeline step of the projection (two steps in flight per thread)

// packed float32 pairs (Blackwell FFMA2): one instruction per two multiply-adds of the inner loops, which are bound by
// instruction issue at the 8 warps per SM their accumulator count allows
__device__ __forceinline__ uint64_t bg_pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void bg_unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t bg_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// part[g][k][f] = sum over the pixels of range g   (KM = 16: background removal; KM = 32: the sketch coefficients of the
// background rSVD, csrc/bgbasis.cu)
template <int KM>
__global__ void __launch_bounds__(kBGThreads)
bg_project_t_kernel(const float* __restrict__ yt, int64_t ld, int64_t d, const float* __restrict__ bg, int k_n, int64_t pix_per_cta,
                    float* __restrict__ part) {
    __shared__ float sb[KM][kBGPix];
    const int64_t f = ((int64_t)blockIdx.x * kBGThreads + threadIdx.x) * 4;
    const int64_t p0 = (int64_t)blockIdx.y * pix_per_cta, p1 = min(d, p0 + pix_per_cta);
    uint64_t acc[KM][2];   // (frames f, f + 1) and (f + 2, f + 3) of component k
#pragma unroll
    for (int k = 0; k < KM; ++k) acc[k][0] = acc[k][1] = 0ull;
    for (int64_t pb = p0; pb < p1; pb += kBGPix) {
        const int n = (int)min((int64_t)kBGPix, p1 - pb);
        __syncthreads();
        for (int i = threadIdx.x; i < KM * kBGPix; i += kBGThreads) {
            const int k = i / kBGPix, q = i - k * kBGPix;
            sb[k][q] = (k < k_n && q < n) ? bg[(int64_t)k * d + pb + q] : 0.f;
        }
        __syncthreads();
        if (f < ld) {
            // software pipeline: the next kBGBatch pixel rows are in flight while the current ones are consumed
            const float* src = yt + pb * ld + f;
            float4 cur[kBGBatch], nxt[kBGBatch];
            auto load = [&](float4 (&y)[kBGBatch], int q0) {
#pragma unroll
                for (int j = 0; j < kBGBatch; ++j)
                    y[j] = q0 + j < n ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)(q0 + j) * ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            load(cur, 0);
            for (int q0 = 0; q0 < n; q0 += kBGBatch) {
                load(nxt, q0 + kBGBatch);
#pragma unroll
                for (int j = 0; j < kBGBatch; ++j) {
                    const uint64_t y01 = bg_pack2(cur[j].x, cur[j].y), y23 = bg_pack2(cur[j].z, cur[j].w);
#pragma unroll
                    for (int k = 0; k < KM; ++k) {
                        const float b = sb[k][min(q0 + j, kBGPix - 1)];
                        const uint64_t bb = bg_pack2(b, b);
                        acc[k][0] = bg_fma2(bb, y01, acc[k][0]);
                        acc[k][1] = bg_fma2(bb, y23, acc[k][1]);
                    }
                }
#pragma unroll
                for (int j = 0; j < kBGBatch; ++j) cur[j] = nxt[j];
            }
        }
    }
    if (f < ld) {
        float* out = part + (int64_t)blockIdx.y * k_n * ld + f;
#pragma unroll
        for (int k = 0; k < KM; ++k)
            if (k < k_n) {
                float4 a;
                bg_unpack2(acc[k][0], a.x, a.y);
                bg_unpack2(acc[k][1], a.z, a.w);
                *reinterpret_cast<float4*>(out + (int64_t)k * ld) = a;
            }
    }
}

__global__ void __launch_bounds__(kBGThreads)
bg_remove_t_kernel(float* __restrict__ yt, int64_t ld, int64_t d, const float* __restrict__ bg, int k_n, const float* __restrict__ vbg,
                   int64_t pix_per_cta) {
    __shared__ float sb[kBGK][kBGPix];
    const int64_t f = ((int64_t)blockIdx.x * kBGThreads + threadIdx.x) * 4;
    const int64_t p0 = (int64_t)blockIdx.y * pix_per_cta, p1 = min(d, p0 + pix_per_cta);
    uint64_t v[kBGK][2];
#pragma unroll
    for (int k = 0; k < kBGK; ++k) {
        const float4 t = (k < k_n && f < ld) ? *reinterpret_cast<const float4*>(vbg + (int64_t)k * ld + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[k][0] = bg_pack2(t.x, t.y);
        v[k][1] = bg_pack2(t.z, t.w);
    }
    for (int64_t pb = p0; pb < p1; pb += kBGPix) {
        const int n = (int)min((int64_t)kBGPix, p1 - pb);
        __syncthreads();
        for (int i = threadIdx.x; i < kBGK * kBGPix; i += kBGThreads) {
            const int k = i / kBGPix, q = i - k * kBGPix;
            sb[k][q] = (k < k_n && q < n) ? bg[(int64_t)k * d + pb + q] : 0.f;
        }
        __syncthreads();
        if (f < ld) {
            // software pipeline as in the projection: the next rows are loaded while the current ones are updated
            float* dst = yt + pb * ld + f;
            float4 cur[kBGBatch], nxt[kBGBatch];
            auto load = [&](float4 (&y)[kBGBatch], int q0) {
#pragma unroll
                for (int j = 0; j < kBGBatch; ++j)
                    if (q0 + j < n) y[j] = *reinterpret_cast<const float4*>(dst + (int64_t)(q0 + j) * ld);
            };
            load(cur, 0);
            for (int q0 = 0; q0 < n; q0 += kBGBatch) {
                load(nxt, q0 + kBGBatch);
#pragma unroll
                for (int j = 0; j < kBGBatch; ++j) {
                    if (q0 + j < n) {
                        uint64_t y01 = bg_pack2(cur[j].x, cur[j].y), y23 = bg_pack2(cur[j].z, cur[j].w);
#pragma unroll
                        for (int k = 0; k < kBGK; ++k) {
                            const float b = -sb[k][q0 + j];
                            const uint64_t bb = bg_pack2(b, b);
                            y01 = bg_fma2(bb, v[k][0], y01);
                            y23 = bg_fma2(bb, v[k][1], y23);
                        }
                        float4 y;
                        bg_unpack2(y01, y.x, y.y);
                        bg_unpack2(y23, y.z, y.w);
                        *reinterpret_cast<float4*>(dst + (int64_t)(q0 + j) * ld) = y;
                    }
                }
#pragma unroll
                for (int j = 0; j < kBGBatch; ++j) cur[j] = nxt[j];
            }
        }
    }
}

}  // namespace pmd

extern "C" int pmd_bg_project_t(const float* yt, int64_t ld, int64_t d, const float* bg, int64_t k, int64_t n_ranges, float* part,
                                void* stream) {
    const char* fn = "pmd_bg_project_t";
    PMD_REQUIRE(yt && bg && part, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && d > 0 && k > 0 && k <= 32 && n_ranges > 0 && n_ranges <= 65535, fn,
                "bad size (ld multiple of 4, 1 <= k <= 32)");
    PMD_REQUIRE(((uintptr_t)yt % 16) == 0 && ((uintptr_t)part % 16) == 0, fn, "operands must be 16-byte aligned");
    const int64_t per = (d + n_ranges - 1) / n_ranges;
    dim3 grid((unsigned)((ld / 4 + pmd::kBGThreads - 1) / pmd::kBGThreads), (unsigned)n_ranges);
    if (k <= 16) pmd::bg_project_t_kernel<16><<<grid, pmd::kBGThreads, 0, (cudaStream_t)stream>>>(yt, ld, d, bg, (int)k, per, part);
    else if (k <= 26) pmd::bg_project_t_kernel<26><<<grid, pmd::kBGThreads, 0, (cudaStream_t)stream>>>(yt, ld, d, bg, (int)k, per, part);
    else pmd::bg_project_t_kernel<32><<<grid, pmd::kBGThreads, 0, (cudaStream_t)stream>>>(yt, ld, d, bg, (int)k, per, part);
    return pmd::check_launch(fn);
}

extern "C" int pmd_bg_remove_t(float* yt, int64_t ld, int64_t d, const float* bg, int64_t k, const float* vbg, void* stream) {
    const char* fn = "pmd_bg_remove_t";
    PMD_REQUIRE(yt && bg && vbg, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && d > 0 && k > 0 && k <= pmd::kBGK, fn, "bad size (ld multiple of 4, 1 <= k <= 16)");
    PMD_REQUIRE(((uintptr_t)yt % 16) == 0 && ((uintptr_t)vbg % 16) == 0, fn, "operands must be 16-byte aligned");
    const int64_t n_ranges = std::min<int64_t>(1024, (d + 255) / 256);
    const int64_t per = (d + n_ranges - 1) / n_ranges;
    dim3 grid((unsigned)((ld / 4 + pmd::kBGThreads - 1) / pmd::kBGThreads), (unsigned)n_ranges);
    pmd::bg_remove_t_kernel<<<grid, pmd::kBGThreads, 0, (cudaStream_t)stream>>>(yt, ld, d, bg, (int)k, vbg, per);
    return pmd::check_launch(fn);
}
