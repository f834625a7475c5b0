// Operand preparation for float32-accurate GEMMs on the tensor cores with ONE TF32 product and ONE bf16 product
// (the scheme of K1 / K7 applied to the library GEMMs of the mixing step, pmd_loader.py:411-412):
//     x = hi + lo,  hi exactly representable in TF32,  lo = x - hi (exact)
//     a b  ~=  a_hi b_hi  (TF32 GEMM, exact products)  +  [a_lo | a_hi] [b_hi ; b_lo]  (bf16 GEMM of twice the depth)
// The correction operands only need bf16: |lo| <= 2^-11 |x|, so rounding hi and lo to 8 bits leaves 2^-19 relative.
// One pass: x is overwritten by hi, the bf16 operand image is written in the concatenated layout the second GEMM wants.
#include <cuda_bf16.h>

#include "common.cuh"

namespace pmd {

__device__ __forceinline__ float split_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);   // round to nearest TF32
}

// mode 0 (right operand, depth = rows):  pair[r][c] = bf16(hi), pair[rows + r][c] = bf16(lo)
// mode 1 (left operand, depth = cols):   pair[r][c] = bf16(lo), pair[r][cols + c] = bf16(hi)
__global__ void __launch_bounds__(256)
split_tf32_bf16_kernel(float* __restrict__ x, int64_t rows, int64_t cols, int64_t ldx, __nv_bfloat16* __restrict__ pair, int64_t ldp,
                       int mode) {
    const int64_t c4 = cols / 4;
    const int64_t total = rows * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / c4, c = (i - r * c4) * 4;
        float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        float4 h = make_float4(split_hi(v.x), split_hi(v.y), split_hi(v.z), split_hi(v.w));
        *reinterpret_cast<float4*>(x + r * ldx + c) = h;
        __nv_bfloat162 h01 = __floats2bfloat162_rn(h.x, h.y), h23 = __floats2bfloat162_rn(h.z, h.w);
        __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - h.x, v.y - h.y), l23 = __floats2bfloat162_rn(v.z - h.z, v.w - h.w);
        uint2 hp, lp;
        hp.x = *reinterpret_cast<uint32_t*>(&h01); hp.y = *reinterpret_cast<uint32_t*>(&h23);
        lp.x = *reinterpret_cast<uint32_t*>(&l01); lp.y = *reinterpret_cast<uint32_t*>(&l23);
        if (mode == 0) {
            *reinterpret_cast<uint2*>(pair + r * ldp + c) = hp;
            *reinterpret_cast<uint2*>(pair + (rows + r) * ldp + c) = lp;
        } else {
            *reinterpret_cast<uint2*>(pair + r * ldp + c) = lp;
            *reinterpret_cast<uint2*>(pair + r * ldp + cols + c) = hp;
        }
    }
}

}  // namespace pmd

extern "C" int pmd_split_tf32_bf16(float* x, int64_t rows, int64_t cols, int64_t ldx, void* pair, int64_t ldp, int mode, void* stream) {
    const char* fn = "pmd_split_tf32_bf16";
    PMD_REQUIRE(x && pair, fn, "null pointer");
    PMD_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && ldx >= cols && ldx % 4 == 0 && ldp % 4 == 0, fn,
                "bad size (cols, ldx, ldp multiples of 4)");
    PMD_REQUIRE(mode == 0 ? ldp >= cols : ldp >= 2 * cols, fn, "pair pitch too small");
    PMD_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)pair % 8) == 0, fn, "operands must be 16-byte (x) / 8-byte (pair) aligned");
    const int64_t total = rows * (cols / 4);
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    pmd::split_tf32_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ldx, (__nv_bfloat16*)pair, ldp, mode);
    return pmd::check_launch(fn);
}
