// Block-stage contractions over the PIXEL-MAJOR init movie  yT[p][f]  (frames contiguous, leading
// dimension ld, a multiple of 4 whose padding columns hold zeros).
//
// The reference fits every overlapping block separately (decomposition.py:790-838 -> single_block_md,
// 235-330); its three t-long contractions per block (295-298, 304-306, 318) are the arithmetic of the
// stage.  Here they are two batched kernels over ALL blocks, both streaming perfectly coalesced
// cp.async tiles (16-byte pieces along frames) through a 3-stage shared-memory pipeline and computing
// with 8x8 register tiles (one LDS.128 per 16 FMAs):
//   block_project_t : out[b][c][f] = sum_q w[b][q][c] * yT[pix(b,q)][f]      (K = block pixels)
//   block_spatial_t : s[b][q][c]   = sum_f yT[pix(b,q)][f] * v[b][c][f]      (K = frames)
// plus the transposing standardisation that produces yT and the pooled / time-averaged sketch input.
#include "common.cuh"

namespace pmd {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------------------------------------
// yT[p][i] = (movie[frames[i]][p] - mean[p]) / stdv[p]     (32 x 32 transposing tiles)
// ------------------------------------------------------------------------------------------------
// A CTA transposes a [128 frames x 128 pixels] tile (kSTTileSets = 4 neighbouring 32-pixel column sets x kSTFrameSets = 4
// sets of 32 frames): a warp reads 4 x 128 contiguous bytes of the same frame back to back and writes 4 x 128 contiguous
// bytes of the same pixel row back to back, so DRAM sees 512-byte bursts on both sides instead of isolated 128-byte lines
// (the pixel rows of the output are ld * 4 bytes apart); 16 independent loads in flight per thread.
constexpr int kSTTileSets = 4;
constexpr int kSTFrameSets = 4;

template <typename T>
__global__ void __launch_bounds__(256)
standardize_frames_t_kernel(const T* __restrict__ movie, int64_t d, const int64_t* __restrict__ frames, int64_t n,
                            const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ out,
                            int64_t ld) {
    extern __shared__ float st_tile[];   // [kSTTileSets][32 * kSTFrameSets][33]
    auto tile = [&](int s, int f, int p) -> float& { return st_tile[(s * (32 * kSTFrameSets) + f) * 33 + p]; };
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * (32 * kSTTileSets), i0 = (int64_t)blockIdx.y * (32 * kSTFrameSets);
    float mu[kSTTileSets], sd[kSTTileSets];
#pragma unroll
    for (int s = 0; s < kSTTileSets; ++s) {
        const int64_t p = p0 + 32 * s + tx;
        mu[s] = p < d ? mean[p] : 0.f;
        sd[s] = p < d ? stdv[p] : 1.f;
    }
#pragma unroll 1
    for (int fs = 0; fs < kSTFrameSets; ++fs) {
        int64_t fr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = i0 + 32 * fs + ty + 8 * j;
            fr[j] = i < n ? frames[i] : -1;
        }
        float raw[kSTTileSets][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // the 4 lines of one frame are requested back to back
#pragma unroll
            for (int s = 0; s < kSTTileSets; ++s) {
                const int64_t p = p0 + 32 * s + tx;
                raw[s][j] = (fr[j] >= 0 && p < d) ? to_f32(movie[fr[j] * d + p]) : 0.f;
            }
        }
#pragma unroll
        for (int s = 0; s < kSTTileSets; ++s) {
            const int64_t p = p0 + 32 * s + tx;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                tile(s, 32 * fs + ty + 8 * j, tx) = (fr[j] >= 0 && p < d) ? (raw[s][j] - mu[s]) / sd[s] : 0.f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < kSTTileSets; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t pp = p0 + 32 * s + ty + 8 * j;
            if (pp >= d) continue;
#pragma unroll
            for (int fs = 0; fs < kSTFrameSets; ++fs) {   // the 4 lines of one pixel row are written back to back
                const int64_t i = i0 + 32 * fs + tx;
                if (i < ld) out[pp * ld + i] = tile(s, 32 * fs + tx, ty + 8 * j);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// pooled + time-averaged block  bta[b][p][tau]  (decomposition.py:192-232 + 283-290)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
block_pool_tavg_t_kernel(const float* __restrict__ yT, int64_t ld, int64_t t, int64_t d2, const int32_t* __restrict__ starts,
                         int bh, int bw, int saf, int taf, float* __restrict__ bta) {
    const int ph = (bh + saf - 1) / saf, pw = (bw + saf - 1) / saf;
    const int lo_h = (ph * saf - bh) / 2, lo_w = (pw * saf - bw) / 2;
    const int P = ph * pw;
    const int64_t tp = t / taf;
    const int64_t b = blockIdx.y;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int64_t tau = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= tp) return;
    for (int p = 0; p < P; ++p) {
        const int pi = p / pw, pj = p - pi * pw;
        const int r0 = max(pi * saf - lo_h, 0), r1 = min(pi * saf - lo_h + saf, bh);
        const int c0 = max(pj * saf - lo_w, 0), c1 = min(pj * saf - lo_w + saf, bw);
        const float cnt = (float)((r1 - r0) * (c1 - c0));
        float acc = 0.f;
        for (int ff = 0; ff < taf; ++ff) {
            const int64_t f = tau * taf + ff;
            float sum = 0.f;
            for (int r = r0; r < r1; ++r)
                for (int c = c0; c < c1; ++c) sum += yT[((int64_t)(i0 + r) * d2 + j0 + c) * ld + f];
            acc += sum / cnt;
        }
        bta[(b * P + p) * tp + tau] = acc / (float)taf;
    }
}

// ------------------------------------------------------------------------------------------------
// pooled block at full time resolution  pooled[b][p][f]  AND its time average  bta[b][p][tau]  in one pass
// (decomposition.py:279 B_ds and 283-290 B_ta): a streaming pooling kernel (thread = frame) + a tiny averaging kernel.
// The pooled tensor lets the first block projection (decomposition.py:295-298, U_ds^T B_ds) contract over the
// P = bpix/4 pooled pixels instead of the bpix full-resolution ones.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
block_pool_full_kernel(const float* __restrict__ yT, int64_t ld, int64_t t, int64_t d2, const int32_t* __restrict__ starts,
                       int bh, int bw, int saf, float* __restrict__ pooled) {
    const int ph = (bh + saf - 1) / saf, pw = (bw + saf - 1) / saf;
    const int lo_h = (ph * saf - bh) / 2, lo_w = (pw * saf - bw) / 2;
    const int P = ph * pw;
    const int64_t b = blockIdx.y;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= ld) return;
    const float* base = yT + ((int64_t)i0 * d2 + j0) * ld + f;
    float* out = pooled + b * P * ld + f;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
        const int pi = p / pw, pj = p - pi * pw;
        const int r0 = max(pi * saf - lo_h, 0), r1 = min(pi * saf - lo_h + saf, bh);
        const int c0 = max(pj * saf - lo_w, 0), c1 = min(pj * saf - lo_w + saf, bw);
        float v = 0.f;
        if (f < t) {
            float sum = 0.f;
            for (int r = r0; r < r1; ++r)
                for (int c = c0; c < c1; ++c) sum += base[((int64_t)r * d2 + c) * ld];
            v = sum / (float)((r1 - r0) * (c1 - c0));
        }
        out[(int64_t)p * ld] = v;
    }
}

// the same with 4 consecutive frames per thread (16-byte loads / stores: 512-byte requests per warp); identical arithmetic
__global__ void __launch_bounds__(256)
block_pool_full4_kernel(const float* __restrict__ yT, int64_t ld, int64_t t, int64_t d2, const int32_t* __restrict__ starts,
                        int bh, int bw, int saf, float* __restrict__ pooled) {
    const int ph = (bh + saf - 1) / saf, pw = (bw + saf - 1) / saf;
    const int lo_h = (ph * saf - bh) / 2, lo_w = (pw * saf - bw) / 2;
    const int P = ph * pw;
    const int64_t b = blockIdx.y;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (f >= ld) return;
    const float* base = yT + ((int64_t)i0 * d2 + j0) * ld + f;
    float* out = pooled + b * P * ld + f;
#pragma unroll 2
    for (int p = 0; p < P; ++p) {
        const int pi = p / pw, pj = p - pi * pw;
        const int r0 = max(pi * saf - lo_h, 0), r1 = min(pi * saf - lo_h + saf, bh);
        const int c0 = max(pj * saf - lo_w, 0), c1 = min(pj * saf - lo_w + saf, bw);
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = r0; r < r1; ++r)
            for (int c = c0; c < c1; ++c) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)r * d2 + c) * ld));
                sum.x += y.x; sum.y += y.y; sum.z += y.z; sum.w += y.w;
            }
        const float cnt = (float)((r1 - r0) * (c1 - c0));
        float4 v;
        v.x = f + 0 < t ? sum.x / cnt : 0.f;
        v.y = f + 1 < t ? sum.y / cnt : 0.f;
        v.z = f + 2 < t ? sum.z / cnt : 0.f;
        v.w = f + 3 < t ? sum.w / cnt : 0.f;
        *reinterpret_cast<float4*>(out + (int64_t)p * ld) = v;
    }
}

// bta[row][tau] = mean of pooled[row][tau*taf .. tau*taf+taf)   (rows = nb * P)
__global__ void __launch_bounds__(256)
block_tavg_kernel(const float* __restrict__ pooled, int64_t ld, int64_t nrows, int64_t tp, int taf, float* __restrict__ bta) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * tp) return;
    const int64_t row = idx / tp, tau = idx - row * tp;
    const float* src = pooled + row * ld + tau * taf;
    float acc = 0.f;
    for (int ff = 0; ff < taf; ++ff) acc += src[ff];
    bta[idx] = acc / (float)taf;
}

// ------------------------------------------------------------------------------------------------
// out[b][c][f] = sum_q w[b][q][c] * yT[pix(b,q)][f]
// CTA = (256 frames, block b, group of NG*8 components); warp g owns components 8g..8g+7, lane l owns
// frames 4l..4l+3 and 128+4l..128+4l+3 of the tile.  K runs over the block's pixels in chunks of <= 16
// pixels of one window row.
// ------------------------------------------------------------------------------------------------
constexpr int kPF = 256;
constexpr int kPK = 16;
constexpr int kPStages = 3;

template <int NG>
__global__ void __launch_bounds__(NG * 32)
block_project_t_kernel(const float* __restrict__ movT, int64_t mbs, int64_t ld, int64_t d2, const int32_t* __restrict__ starts,
                       int bh, int bw, const float* __restrict__ w, int r, int rp, float* __restrict__ out, int64_t ldo) {
    extern __shared__ __align__(16) float psm[];
    constexpr int NC = NG * 8;
    constexpr int STAGE = kPK * kPF + kPK * NC;
    constexpr int NT = NG * 32;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const int64_t b = blockIdx.y;
    const int64_t f0 = (int64_t)blockIdx.x * kPF;
    const int cbase = blockIdx.z * NC;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const float* mv = movT + b * mbs;
    const float* wb = w + b * (int64_t)(bh * bw) * rp;
    const int bpix = bh * bw;
    const int nchunks = (bpix + kPK - 1) / kPK;   // K chunks of 16 consecutive block pixels (row boundaries ignored)
    const int rp4 = rp / 4;

    auto issue = [&](int ch, int st) {
        float* xs = psm + st * STAGE;
        float* ws = xs + kPK * kPF;
        const int q0 = ch * kPK;
        const int kc = min(kPK, bpix - q0);
        for (int idx = tid; idx < kc * (kPF / 4); idx += NT) {
            const int k = idx / (kPF / 4), c4 = idx - k * (kPF / 4);
            const int q = q0 + k;
            const int qi = q / bw, qj = q - qi * bw;
            const int64_t pix = (int64_t)(i0 + qi) * d2 + j0 + qj;
            const int64_t f = f0 + 4 * c4;
            float* dst = xs + k * kPF + 4 * c4;
            if (f < ld) cp_async16(dst, mv + pix * ld + f);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int idx = tid; idx < kc * (NC / 4); idx += NT) {
            const int k = idx / (NC / 4), c4 = idx - k * (NC / 4);
            float* dst = ws + k * NC + 4 * c4;
            const int cg4 = cbase / 4 + c4;
            if (cg4 < rp4) cp_async16(dst, wb + (int64_t)(q0 + k) * rp + 4 * cg4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < kPStages - 1; ++s) {
        if (s < nchunks) issue(s, s);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<kPStages - 2>();
        __syncthreads();
        const int nx = ch + kPStages - 1;
        if (nx < nchunks) issue(nx, nx % kPStages);
        cp_async_commit();
        const float* xs = psm + (ch % kPStages) * STAGE;
        const float* ws = xs + kPK * kPF;
        const int kc = min(kPK, bpix - ch * kPK);
#pragma unroll 4
        for (int k = 0; k < kc; ++k) {
            const float4 xa = *reinterpret_cast<const float4*>(xs + k * kPF + 4 * lane);
            const float4 xb = *reinterpret_cast<const float4*>(xs + k * kPF + 128 + 4 * lane);
            const float4 w0 = *reinterpret_cast<const float4*>(ws + k * NC + 8 * g);
            const float4 w1 = *reinterpret_cast<const float4*>(ws + k * NC + 8 * g + 4);
            const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
    }
    const bool vec = (ldo % 4) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cbase + 8 * g + i;
        if (c >= r) continue;
        float* o = out + (b * r + c) * ldo;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t f = f0 + 128 * h + 4 * lane;
            if (vec && f + 3 < ldo) {
                *reinterpret_cast<float4*>(o + f) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (f + j < ldo) o[f + j] = acc[i][4 * h + j];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// s[b][q][c] = sum_f yT[pix(b,q)][f] * v[b][c][f]
// CTA = (pixel tile of PT = 8*QT pixels, block b); thread (qt, g) owns pixels qt + QT*i (i < 8) and
// components 8g..8g+7.  K runs over frames in chunks of 32; both operands are read with LDS.128 along
// frames (row stride 36 floats: conflict free).
// ------------------------------------------------------------------------------------------------
constexpr int kSF = 32;
constexpr int kSLD = 36;

template <int NG>
__global__ void __launch_bounds__(384)
block_spatial_t_kernel(const float* __restrict__ movT, int64_t mbs, int64_t ld, int64_t d2, const int32_t* __restrict__ starts,
                       int bh, int bw, const float* __restrict__ v, int64_t ldv, int r, int rp, float* __restrict__ s, int PT,
                       int QT, int nstages) {
    extern __shared__ __align__(16) float ssm[];
    constexpr int NC = NG * 8;
    const int NT = QT * NG;
    const int tid = threadIdx.x;
    const int qt = tid % QT, g = tid / QT;
    const int64_t b = blockIdx.y;
    const int ptile = blockIdx.x;
    const int bpix = bh * bw;
    const int i0 = starts[2 * b], j0 = starts[2 * b + 1];
    const float* mv = movT + b * mbs;
    const float* vb = v + b * (int64_t)r * ldv;
    const int stage_floats = (PT + NC) * kSLD;
    int* pixoff = reinterpret_cast<int*>(ssm + (size_t)nstages * stage_floats);
    for (int pl = tid; pl < PT; pl += NT) {
        const int q = ptile * PT + pl;
        int off = -1;
        if (q < bpix) {
            const int qi = q / bw, qj = q - qi * bw;
            off = (i0 + qi) * (int)d2 + j0 + qj;
        }
        pixoff[pl] = off;
    }
    __syncthreads();
    const int nchunks = (int)((ld + kSF - 1) / kSF);

    auto issue = [&](int ch, int st) {
        float* xs = ssm + (size_t)st * stage_floats;
        float* vs = xs + PT * kSLD;
        const int64_t fc = (int64_t)ch * kSF;
        for (int idx = tid; idx < PT * (kSF / 4); idx += NT) {
            const int pl = idx >> 3, c4 = idx & 7;
            const int off = pixoff[pl];
            const int64_t f = fc + 4 * c4;
            float* dst = xs + pl * kSLD + 4 * c4;
            if (off >= 0 && f < ld) cp_async16(dst, mv + (int64_t)off * ld + f);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int idx = tid; idx < NC * (kSF / 4); idx += NT) {
            const int c = idx >> 3, c4 = idx & 7;
            const int64_t f = fc + 4 * c4;
            float* dst = vs + c * kSLD + 4 * c4;
            if (c < r && f < ldv) cp_async16(dst, vb + (int64_t)c * ldv + f);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int st = 0; st < nstages - 1; ++st) {
        if (st < nchunks) issue(st, st);
        cp_async_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        if (nstages == 3) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();
        const int nx = ch + nstages - 1;
        if (nx < nchunks) issue(nx, nx % nstages);
        cp_async_commit();
        const float* xs = ssm + (size_t)(ch % nstages) * stage_floats;
        const float* vs = xs + PT * kSLD;
#pragma unroll 2
        for (int k4 = 0; k4 < kSF / 4; ++k4) {
            float4 vv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) vv[j] = *reinterpret_cast<const float4*>(vs + (8 * g + j) * kSLD + 4 * k4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + (qt + QT * i) * kSLD + 4 * k4);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a = acc[i][j];
                    a = fmaf(xv.x, vv[j].x, a);
                    a = fmaf(xv.y, vv[j].y, a);
                    a = fmaf(xv.z, vv[j].z, a);
                    a = fmaf(xv.w, vv[j].w, a);
                    acc[i][j] = a;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int q = ptile * PT + qt + QT * i;
        if (qt + QT * i >= PT || q >= bpix) continue;
        float* o = s + ((int64_t)b * bpix + q) * rp;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 8 * g + 4 * h;
            if (c < rp) *reinterpret_cast<float4*>(o + c) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
        }
    }
}

}  // namespace pmd

// =================================================================================================
extern "C" int pmd_standardize_frames_t(const void* movie, int dtype, int64_t d, const int64_t* frames, int64_t n_frames,
                                        const float* mean, const float* stdv, float* out, int64_t ld, void* stream) {
    const char* fn = "pmd_standardize_frames_t";
    PMD_REQUIRE(movie && frames && mean && stdv && out, fn, "null pointer");
    PMD_REQUIRE(d > 0 && n_frames > 0 && ld >= n_frames, fn, "bad size");
    const int64_t gy = (ld + 32 * pmd::kSTFrameSets - 1) / (32 * pmd::kSTFrameSets);
    PMD_REQUIRE(gy <= 65535, fn, "more than 65535*128 frames per call");
    dim3 grid((unsigned)((d + 32 * pmd::kSTTileSets - 1) / (32 * pmd::kSTTileSets)), (unsigned)gy);
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = pmd::kSTTileSets * 32 * pmd::kSTFrameSets * 33 * (int)sizeof(float);
    PMD_DISPATCH_DTYPE(dtype, fn, {
        cudaError_t e = cudaFuncSetAttribute(pmd::standardize_frames_t_kernel<scalar_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
        pmd::standardize_frames_t_kernel<scalar_t><<<grid, 256, smem, st>>>((const scalar_t*)movie, d, frames, n_frames, mean, stdv,
                                                                            out, ld);
    });
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_pool_tavg(const float* yt, int64_t ld, int64_t t, int64_t d2, const int32_t* starts, int64_t nb,
                                   int64_t bh, int64_t bw, int64_t saf, int64_t taf, float* bta, void* stream) {
    const char* fn = "pmd_block_pool_tavg";
    PMD_REQUIRE(yt && starts && bta, fn, "null pointer");
    PMD_REQUIRE(t > 0 && ld >= t && nb > 0 && nb <= 65535 && bh > 0 && bw > 0 && saf > 0 && taf > 0 && t / taf > 0, fn, "bad size");
    const int64_t tp = t / taf;
    dim3 grid((unsigned)((tp + 127) / 128), (unsigned)nb);
    pmd::block_pool_tavg_t_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(yt, ld, t, d2, starts, (int)bh, (int)bw, (int)saf,
                                                                         (int)taf, bta);
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_pool_full(const float* yt, int64_t ld, int64_t t, int64_t d2, const int32_t* starts, int64_t nb,
                                   int64_t bh, int64_t bw, int64_t saf, int64_t taf, float* pooled, float* bta, void* stream) {
    const char* fn = "pmd_block_pool_full";
    PMD_REQUIRE(yt && starts && pooled && bta, fn, "null pointer");
    PMD_REQUIRE(t > 0 && ld >= t && nb > 0 && nb <= 65535 && bh > 0 && bw > 0 && saf > 0 && taf > 0 && t % taf == 0, fn,
                "bad size (t a multiple of taf)");
    const int64_t P = ((bh + saf - 1) / saf) * ((bw + saf - 1) / saf);
    const int64_t tp = t / taf;
    cudaStream_t st = (cudaStream_t)stream;
    if (ld % 4 == 0 && ((uintptr_t)yt % 16) == 0 && ((uintptr_t)pooled % 16) == 0) {
        dim3 grid((unsigned)((ld / 4 + 255) / 256), (unsigned)nb);
        pmd::block_pool_full4_kernel<<<grid, 256, 0, st>>>(yt, ld, t, d2, starts, (int)bh, (int)bw, (int)saf, pooled);
    } else {
        dim3 grid((unsigned)((ld + 255) / 256), (unsigned)nb);
        pmd::block_pool_full_kernel<<<grid, 256, 0, st>>>(yt, ld, t, d2, starts, (int)bh, (int)bw, (int)saf, pooled);
    }
    const int64_t total = nb * P * tp;
    pmd::block_tavg_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pooled, ld, nb * P, tp, (int)taf, bta);
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_project(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                                 const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w, int64_t r,
                                 int64_t rp, float* out, int64_t ldo, void* stream) {
    const char* fn = "pmd_block_project";
    PMD_REQUIRE(movie_t && starts && w && out, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && ldo > 0 && ldo <= ld && nb > 0 && nb <= 65535 && r > 0 && rp >= r && rp % 4 == 0, fn,
                "bad size (ld multiple of 4, ldo <= ld, rp multiple of 4 and >= r)");
    PMD_REQUIRE(((uintptr_t)movie_t % 16) == 0 && ((uintptr_t)w % 16) == 0 && (movie_batch_stride % 4) == 0, fn,
                "operands must be 16-byte aligned");
    const int groups = (int)((r + 7) / 8);
    const int ng = groups >= 8 ? 8 : groups;
    const int gz = (groups + ng - 1) / ng;
    dim3 grid((unsigned)((ldo + pmd::kPF - 1) / pmd::kPF), (unsigned)nb, (unsigned)gz);
    cudaStream_t st = (cudaStream_t)stream;
#define PMD_LAUNCH_PROJ(NG)                                                                                              \
    case NG: {                                                                                                           \
        auto k = pmd::block_project_t_kernel<NG>;                                                                        \
        const size_t smem = (size_t)pmd::kPStages * (pmd::kPK * pmd::kPF + pmd::kPK * NG * 8) * sizeof(float);           \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }         \
        k<<<grid, NG * 32, smem, st>>>(movie_t, movie_batch_stride, ld, d2, starts, (int)bh, (int)bw, w, (int)r, (int)rp, \
                                       out, ldo);                                                                        \
        break;                                                                                                           \
    }
    switch (ng) {
        PMD_LAUNCH_PROJ(1)
        PMD_LAUNCH_PROJ(2)
        PMD_LAUNCH_PROJ(3)
        PMD_LAUNCH_PROJ(4)
        PMD_LAUNCH_PROJ(5)
        PMD_LAUNCH_PROJ(6)
        PMD_LAUNCH_PROJ(7)
        PMD_LAUNCH_PROJ(8)
    }
#undef PMD_LAUNCH_PROJ
    return pmd::check_launch(fn);
}

extern "C" int pmd_block_spatial(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                                 const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* v, int64_t ldv,
                                 int64_t r, int64_t rp, float* s, void* stream) {
    const char* fn = "pmd_block_spatial";
    PMD_REQUIRE(movie_t && starts && v && s, fn, "null pointer");
    PMD_REQUIRE(ld > 0 && ld % 4 == 0 && ldv > 0 && ldv % 4 == 0 && ldv <= ld && nb > 0 && nb <= 65535 && r > 0 && rp >= r &&
                    rp % 4 == 0 && rp <= 64,
                fn, "bad size (ld, ldv multiples of 4, ldv <= ld, rp multiple of 4, r <= rp <= 64)");
    PMD_REQUIRE(((uintptr_t)movie_t % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)s % 16) == 0 &&
                    (movie_batch_stride % 4) == 0,
                fn, "operands must be 16-byte aligned");
    const int bpix = (int)(bh * bw);
    const int ng = (int)((rp + 7) / 8);
    const int max_qt = std::min(384 / ng, 88);                     // threads <= 384, two stages fit shared memory
    const int ntiles = (bpix + 8 * max_qt - 1) / (8 * max_qt);
    const int QT = ((bpix + ntiles - 1) / ntiles + 7) / 8;
    const int PT = 8 * QT;
    const size_t stage = (size_t)(PT + ng * 8) * pmd::kSLD * sizeof(float);
    const int nstages = (3 * stage + PT * sizeof(int) <= 220 * 1024) ? 3 : 2;
    const size_t smem = nstages * stage + PT * sizeof(int);
    PMD_REQUIRE(smem <= 227 * 1024, fn, "tile does not fit shared memory");
    dim3 grid((unsigned)ntiles, (unsigned)nb);
    cudaStream_t st = (cudaStream_t)stream;
#define PMD_LAUNCH_SPAT(NG)                                                                                              \
    case NG: {                                                                                                           \
        auto k = pmd::block_spatial_t_kernel<NG>;                                                                        \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }         \
        k<<<grid, QT * NG, smem, st>>>(movie_t, movie_batch_stride, ld, d2, starts, (int)bh, (int)bw, v, ldv, (int)r,     \
                                       (int)rp, s, PT, QT, nstages);                                                     \
        break;                                                                                                           \
    }
    switch (ng) {
        PMD_LAUNCH_SPAT(1)
        PMD_LAUNCH_SPAT(2)
        PMD_LAUNCH_SPAT(3)
        PMD_LAUNCH_SPAT(4)
        PMD_LAUNCH_SPAT(5)
        PMD_LAUNCH_SPAT(6)
        PMD_LAUNCH_SPAT(7)
        PMD_LAUNCH_SPAT(8)
    }
#undef PMD_LAUNCH_SPAT
    return pmd::check_launch(fn);
}
