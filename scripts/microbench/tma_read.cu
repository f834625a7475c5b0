// Calibration for K7: how fast can the SMs pull a frame-major movie (T frames x d1 rows x d2 pixels, float32) into
// shared memory when every CTA walks DOWN the image rows of a column strip for a group of frames?
//   request shape = [F frames x W pixels] per image row, fetched as 2-D TMA boxes (cp.async.bulk.tensor.2d) into a ring
//   of STAGES shared-memory stages; one thread issues, one thread releases (the data is not consumed).
// Prints achieved GB/s for several (W, F, CTA order) combinations so that the strip width / frame group of the
// projection kernel can be chosen from measurements.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_read tma_read.cu   (no -lcuda: driver entry point at run time)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);   \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "W_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra W_DONE;\n\t"
        "bra W_WAIT;\n\t"
        "W_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

constexpr int kMaxStages = 12;

// order 0: blockIdx -> frame group fastest (CTAs of one strip neighbours); order 1: strip fastest (the strips of one
// frame group neighbours: together they request whole image rows)
__global__ void __launch_bounds__(64, 1)
tma_read_kernel(const __grid_constant__ CUtensorMap tm, int box_w, int boxes_per_row, int F, int d1, int d2, int n_strips, int n_fg,
                int order, int stages, int stage_bytes, int row_lo, int row_hi) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages];
    const uint32_t sbase = (smem_u32(sm) + 1023u) & ~1023u;
    int strip, fg;
    if (order == 0) {
        fg = blockIdx.x % n_fg;
        strip = blockIdx.x / n_fg;
    } else {
        strip = blockIdx.x % n_strips;
        fg = blockIdx.x / n_strips;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::);
    }
    __syncthreads();
    const int n_rows = row_hi - row_lo;
    row_lo += blockIdx.y * n_rows;
    if (threadIdx.x == 0) {
        for (int r = 0; r < n_rows; ++r) {
            const int s = r % stages;
            if (r >= stages) mbar_wait(smem_u32(&empty[s]), ((r / stages) - 1) & 1);
            const uint32_t bar = smem_u32(&full[s]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"((uint32_t)stage_bytes) : "memory");
            for (int b = 0; b < boxes_per_row; ++b) {
                const int x = (row_lo + r) * d2 + strip * box_w * boxes_per_row + b * box_w;
                const int y = fg * F;
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                        sbase + s * stage_bytes + b * (stage_bytes / boxes_per_row)),
                    "l"(&tm), "r"(x), "r"(y), "r"(bar)
                    : "memory");
            }
        }
    } else if (threadIdx.x == 32) {
        for (int r = 0; r < n_rows; ++r) {
            const int s = r % stages;
            mbar_wait(smem_u32(&full[s]), (r / stages) & 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[s])) : "memory");
        }
    }
}

// plain vector loads for comparison: a warp reads W pixels of one frame per request (W*4/16 lanes active), 8 requests in flight
__global__ void __launch_bounds__(512, 1)
ldg_read_kernel(const float* __restrict__ movie, int64_t d, int W, int F, int d1, int d2, int n_strips, int n_fg, int order,
                float* __restrict__ sink) {
    int strip, fg;
    if (order == 0) {
        fg = blockIdx.x % n_fg;
        strip = blockIdx.x / n_fg;
    } else {
        strip = blockIdx.x % n_strips;
        fg = blockIdx.x / n_strips;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lanes_per_frame = W / 4;                      // W <= 128
    const int frames_per_req = 32 / lanes_per_frame;
    const int fl = lane / lanes_per_frame, c = lane % lanes_per_frame;
    float acc = 0.f;
    const int per_warp = F / 16;                            // frames per warp
    for (int r = blockIdx.y * (d1 / 4); r < (blockIdx.y + 1) * (d1 / 4); ++r) {
        const float* base = movie + (int64_t)(fg * F + warp * per_warp + fl) * d + (int64_t)r * d2 + strip * W + 4 * c;
#pragma unroll 8
        for (int q = 0; q < per_warp; q += frames_per_req) {
            float4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "l"(base + (int64_t)q * d));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8192, d1 = 512, d2 = 512;
    const int64_t d = (int64_t)d1 * d2;
    float* movie;
    CK(cudaMalloc(&movie, (size_t)T * d * 4));
    CK(cudaMemset(movie, 0, (size_t)T * d * 4));
    float* sink;
    CK(cudaMalloc(&sink, 4));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn encode = (EncodeFn)fn;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double bytes = (double)T * d * 4;
    CK(cudaFuncSetAttribute(tma_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg { int W, F, order, swz; };
    std::vector<Cfg> cfgs;
    for (int order = 0; order < 2; ++order)
        for (int W : {32, 64, 128, 256, 512})
            for (int kb : {16, 32}) cfgs.push_back(Cfg{W, kb * 1024 / (W * 4), order, 0});
    cfgs.push_back(Cfg{32, 128, 0, 1});   // the operand shape of the projection kernel: 128 frames x 128 bytes, SWIZZLE_128B
    cfgs.push_back(Cfg{32, 128, 1, 1});
    cfgs.push_back(Cfg{32, 256, 1, 1});
    for (const Cfg& c : cfgs) {
        const int box_w = c.W > 256 ? 256 : c.W, boxes = c.W / box_w;
        const int Fb = c.F > 256 ? 256 : c.F;
        if (Fb != c.F) continue;
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)T};
        cuuint64_t gstr[1] = {(cuuint64_t)d * 4};
        cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)c.F};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, movie, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            c.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("encode failed %d for W %d F %d\n", (int)r, c.W, c.F);
            continue;
        }
        const int stage_bytes = c.W * 4 * c.F;
        const int stages = 192 * 1024 / stage_bytes > kMaxStages ? kMaxStages : 192 * 1024 / stage_bytes;
        const int n_strips = d2 / c.W, n_fg = T / c.F;
        const int grid = n_strips * n_fg;
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            tma_read_kernel<<<dim3(grid, 4), 64, stages * stage_bytes + 1024>>>(tm, box_w, boxes, c.F, d1, d2, n_strips, n_fg, c.order, stages,
                                                                        stage_bytes, 0, d1 / 4);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("tma  W %4d px (%5d B) F %4d stage %3d KB x %2d  order %s swz %d  grid %6d : %8.3f ms  %7.1f GB/s\n", c.W, c.W * 4, c.F,
               stage_bytes / 1024, stages, c.order ? "strip-fastest" : "frame-fastest", c.swz, grid, best, bytes / best / 1e6);
        fflush(stdout);
    }
    // persistent variant: 148 CTAs (one per SM) is what a 1-CTA/SM kernel sees per wave; here grid = all items, 1 CTA/SM by smem
    for (int order = 0; order < 2; ++order)
        for (int W : {32, 64, 128}) {
            const int F = 512;
            const int n_strips = d2 / W, n_fg = T / F;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaEventRecord(e0));
                ldg_read_kernel<<<dim3(n_strips * n_fg, 4), 512>>>(movie, d, W, F, d1, d2, n_strips, n_fg, order, sink);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            printf("ldg  W %4d px F %4d order %s grid %6d : %8.3f ms  %7.1f GB/s\n", W, F, order ? "strip-fastest" : "frame-fastest",
                   n_strips * n_fg, best, bytes / best / 1e6);
            fflush(stdout);
        }
    return 0;
}
