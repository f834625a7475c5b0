// Frame-major gather + standardisation of frames (pmd_loader.py:293-298, 374-377); the stats pass itself lives
// in stats_fft.cu.
#include "common.cuh"

namespace pmd {

template <typename T>
__global__ void standardize_frames_kernel(const T* __restrict__ movie, int64_t d, const int64_t* __restrict__ frames,
                                          const float* __restrict__ mean, const float* __restrict__ stdv,
                                          float* __restrict__ out) {
    const int64_t i = blockIdx.y;
    const int64_t f = frames[i];
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < d; p += (int64_t)gridDim.x * blockDim.x) {
        out[i * d + p] = (to_f32(movie[f * d + p]) - mean[p]) / stdv[p];
    }
}

}  // namespace pmd

extern "C" int pmd_standardize_frames(const void* movie, int dtype, int64_t d, const int64_t* frames, int64_t n_frames,
                                      const float* mean, const float* stdv, float* out, void* stream) {
    const char* fn = "pmd_standardize_frames";
    PMD_REQUIRE(movie && frames && mean && stdv && out, fn, "null pointer");
    PMD_REQUIRE(d > 0 && n_frames >= 0, fn, "bad size");
    if (n_frames == 0) return 0;
    PMD_REQUIRE(n_frames <= 65535, fn, "more than 65535 frames per call");
    dim3 grid((unsigned)std::min<int64_t>((d + 255) / 256, 1024), (unsigned)n_frames);
    cudaStream_t st = (cudaStream_t)stream;
    PMD_DISPATCH_DTYPE(dtype, fn, {
        pmd::standardize_frames_kernel<scalar_t><<<grid, 256, 0, st>>>((const scalar_t*)movie, d, frames, mean, stdv, out);
    });
    return pmd::check_launch(fn);
}
