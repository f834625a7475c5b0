// Shared helpers for libpmd_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/pmd_sm100.h"

namespace pmd {

void set_error(const std::string& msg);

inline int fail_arg(const char* fn, const char* what) {
    set_error(std::string(fn) + ": invalid argument: " + what);
    return -1;
}

inline int check_launch(const char* fn) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error(std::string(fn) + ": " + cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

#define PMD_REQUIRE(cond, fn, what) \
    do {                            \
        if (!(cond)) return pmd::fail_arg(fn, what); \
    } while (0)

template <typename T>
__device__ __forceinline__ float to_f32(T v) {
    return (float)v;
}

// dispatch a templated launcher on the movie element type
#define PMD_DISPATCH_DTYPE(dtype, fn, ...)                                      \
    switch (dtype) {                                                            \
        case PMD_F32: { using scalar_t = float; __VA_ARGS__; break; }           \
        case PMD_U16: { using scalar_t = uint16_t; __VA_ARGS__; break; }        \
        case PMD_I16: { using scalar_t = int16_t; __VA_ARGS__; break; }         \
        case PMD_U8:  { using scalar_t = uint8_t; __VA_ARGS__; break; }         \
        case PMD_F64: { using scalar_t = double; __VA_ARGS__; break; }          \
        case PMD_I32: { using scalar_t = int32_t; __VA_ARGS__; break; }         \
        default: return pmd::fail_arg(fn, "unknown dtype");                     \
    }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pmd
