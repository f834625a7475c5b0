// FFMA vs FFMA2 issue-rate microbenchmark (B200). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma ffma.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { acc[i].x = fmaf(acc[i].x, x.x, y.x); acc[i].y = fmaf(acc[i].y, x.y, y.y); }
                else acc[i] = __ffma2_rn(acc[i], x, y);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000; const int blocks = 148 * 8;
    for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<blocks, 256>>>(out, 0.999f, 0.001f, iters); else k<1><<<blocks, 256>>>(out, 0.999f, 0.001f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)blocks * 256 * iters * 8 * 8 * 2;
        printf("mode %s: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.965GHz)\n", mode ? "FFMA2" : "FFMA ", ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
