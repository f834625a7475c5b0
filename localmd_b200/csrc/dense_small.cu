// Small dense building blocks used by every per-block factorisation: batched Gram matrices with
// float64 accumulation and a batched cyclic-Jacobi symmetric eigensolver in float64.
//
// Why float64 Gram + Jacobi instead of a float32 Householder SVD: the reference calls LAPACK-class
// SVD/QR on r x t (50 x 5000) matrices per block (decomposition.py:64,66,301,315,319).  Forming the
// Gram matrix squares the condition number, which in float32 would corrupt exactly the weak trailing
// components that decide the per-block rank; products of float32 values are exact in float64, so a
// float64 Gram followed by Jacobi (relative-accuracy stopping rule) resolves singular vectors of
// matrices with condition numbers up to ~1e7 better than a float32 backward-stable SVD would.
#include "common.cuh"

namespace pmd {

constexpr int kGramMT = 64;       // inner-dimension tile
constexpr int kGramThreads = 256;
constexpr int kGramMaxTiles = 7;  // ceil(56*57/2 / 256) for n <= 112

// C[b] += A_b[:, m0:m1] A_b[:, m0:m1]^T, fp32 in, fp64 accumulate, 2x2 register tiles over the
// upper triangle, atomically added into C (both triangles).
__global__ void __launch_bounds__(kGramThreads)
gram_f64_kernel(const float* __restrict__ a, int n, int64_t m_len, int64_t bs, int64_t rs, int64_t is,
                int64_t m_per_cta, double* __restrict__ c) {
    extern __shared__ double gsm[];  // [2*nt][kGramMT+1]
    const int nt = (n + 1) / 2;
    const int ld = kGramMT + 1;
    const int ntiles = nt * (nt + 1) / 2;
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    const float* ab = a + b * bs;
    const int64_t m_begin = (int64_t)blockIdx.x * m_per_cta;
    const int64_t m_end = min(m_len, m_begin + m_per_cta);

    int ti[kGramMaxTiles], tj[kGramMaxTiles];
    double acc[kGramMaxTiles][4];
#pragma unroll
    for (int k = 0; k < kGramMaxTiles; ++k) {
        int id = tid + k * kGramThreads;
        ti[k] = -1;
        tj[k] = 0;
        if (id < ntiles) {
            int row = 0, rem = id;
            while (rem >= nt - row) { rem -= nt - row; ++row; }
            ti[k] = row;
            tj[k] = row + rem;
        }
        acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.0;
    }
    // zero the padding row once (odd n)
    if (n & 1) for (int mm = tid; mm < ld; mm += kGramThreads) gsm[(2 * nt - 1) * ld + mm] = 0.0;

    for (int64_t m0 = m_begin; m0 < m_end; m0 += kGramMT) {
        if (is == 1) {
            for (int idx = tid; idx < n * kGramMT; idx += kGramThreads) {
                const int i = idx / kGramMT, mm = idx % kGramMT;
                const int64_t m = m0 + mm;
                gsm[i * ld + mm] = m < m_end ? (double)ab[(int64_t)i * rs + m] : 0.0;
            }
        } else {
            for (int idx = tid; idx < n * kGramMT; idx += kGramThreads) {
                const int mm = idx / n, i = idx % n;
                const int64_t m = m0 + mm;
                gsm[i * ld + mm] = m < m_end ? (double)ab[(int64_t)i * rs + m * is] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kGramMaxTiles; ++k) {
            if (ti[k] >= 0) {
                const double* r0 = gsm + (2 * ti[k]) * ld;
                const double* r1 = r0 + ld;
                const double* q0 = gsm + (2 * tj[k]) * ld;
                const double* q1 = q0 + ld;
                double a00 = acc[k][0], a01 = acc[k][1], a10 = acc[k][2], a11 = acc[k][3];
#pragma unroll 8
                for (int mm = 0; mm < kGramMT; ++mm) {
                    const double x0 = r0[mm], x1 = r1[mm], y0 = q0[mm], y1 = q1[mm];
                    a00 = fma(x0, y0, a00);
                    a01 = fma(x0, y1, a01);
                    a10 = fma(x1, y0, a10);
                    a11 = fma(x1, y1, a11);
                }
                acc[k][0] = a00; acc[k][1] = a01; acc[k][2] = a10; acc[k][3] = a11;
            }
        }
        __syncthreads();
    }
    double* cb = c + b * (int64_t)n * n;
#pragma unroll
    for (int k = 0; k < kGramMaxTiles; ++k) {
        if (ti[k] < 0) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = 2 * ti[k] + (e >> 1), j = 2 * tj[k] + (e & 1);
            if (i >= n || j >= n) continue;
            atomicAdd(&cb[(int64_t)i * n + j], acc[k][e]);
            if (ti[k] != tj[k]) atomicAdd(&cb[(int64_t)j * n + i], acc[k][e]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same Gram matrices on the FP64 tensor cores (mma.sync m8n8k4, DMMA), for rows that are contiguous in memory
// (inner stride 1): C[b] += A_b[:, m0:m1] A_b[:, m0:m1]^T.  float32 inputs are exact in float64, every product is exact
// and the accumulation is IEEE float64, as in the FMA kernel above; only the summation order differs.
// A CTA stages [n rows][kDGK values] float32 tiles in shared memory (row pitch = 4 mod 32 words: the 8 x 4 fragment
// loads of a warp are bank-conflict free); warp w owns the upper-triangle 8 x 8 output tiles w, w + 8, ...
// ------------------------------------------------------------------------------------------------
constexpr int kDGK = 64;                 // inner-dimension values per shared-memory tile
constexpr int kDGLd = kDGK + 4;          // row pitch in floats
constexpr int kDGThreads = 256;
// kDGMaxTiles = upper-triangle 8 x 8 tiles per warp: ceil(14 * 15 / 2 / 8) = 14 for n <= 112, 4 for n <= 56 (the block
// stage's n = 50: 28 tiles on 8 warps) -- the small variant needs 60 instead of 125 registers, which lifts the
// register-limited residency from 2 to 4 CTAs per SM (ncu: tensor pipe 44 % active, long-scoreboard stalls on the tile loads)
template <int kDGMaxTiles>
__global__ void __launch_bounds__(kDGThreads, kDGMaxTiles <= 4 ? 4 : 2)
gram_dmma_kernel(const float* __restrict__ a, int n, int64_t m_len, int64_t bs, int64_t rs, int64_t m_per_cta,
                 double* __restrict__ c) {
    extern __shared__ float dsm[];       // [8 * nt][kDGLd]
    const int nt = (n + 7) / 8;
    const int ntiles = nt * (nt + 1) / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b = blockIdx.y;
    const float* ab = a + b * bs;
    const int64_t m_begin = (int64_t)blockIdx.x * m_per_cta;
    const int64_t m_end = min(m_len, m_begin + m_per_cta);
    int ti[kDGMaxTiles], tj[kDGMaxTiles];
    double acc[kDGMaxTiles][2];
#pragma unroll
    for (int k = 0; k < kDGMaxTiles; ++k) {
        const int id = warp + 8 * k;
        ti[k] = -1;
        tj[k] = 0;
        if (id < ntiles) {
            int row = 0, rem = id;
            while (rem >= nt - row) { rem -= nt - row; ++row; }
            ti[k] = row;
            tj[k] = row + rem;
        }
        acc[k][0] = acc[k][1] = 0.0;
    }
    const int fr = lane >> 2, fc = lane & 3;   // fragment element of this lane: row fr (0..7), k offset fc (0..3)
    for (int64_t m0 = m_begin; m0 < m_end; m0 += kDGK) {
        __syncthreads();
        for (int idx = tid; idx < 8 * nt * (kDGK / 4); idx += kDGThreads) {
            const int i = idx / (kDGK / 4), q = idx - i * (kDGK / 4);
            const int64_t m = m0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n) {
                const float* src = ab + (int64_t)i * rs + m;
                if (m + 3 < m_end && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
                    v = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                    if (m < m_end) v.x = src[0];
                    if (m + 1 < m_end) v.y = src[1];
                    if (m + 2 < m_end) v.z = src[2];
                    if (m + 3 < m_end) v.w = src[3];
                }
            }
            *reinterpret_cast<float4*>(dsm + i * kDGLd + 4 * q) = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kDGMaxTiles; ++k) {
            if (ti[k] < 0) continue;
            const float* pa = dsm + (8 * ti[k] + fr) * kDGLd + fc;
            const float* pb = dsm + (8 * tj[k] + fr) * kDGLd + fc;
            double c0 = acc[k][0], c1 = acc[k][1];
#pragma unroll
            for (int kk = 0; kk < kDGK; kk += 4) {
                const double av = (double)pa[kk], bv = (double)pb[kk];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                             : "+d"(c0), "+d"(c1)
                             : "d"(av), "d"(bv));
            }
            acc[k][0] = c0;
            acc[k][1] = c1;
        }
    }
    // C fragment: lane holds C[fr][2 fc], C[fr][2 fc + 1] of its tile
    double* cb = c + b * (int64_t)n * n;
#pragma unroll
    for (int k = 0; k < kDGMaxTiles; ++k) {
        if (ti[k] < 0) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int i = 8 * ti[k] + fr, j = 8 * tj[k] + 2 * fc + e;
            if (i >= n || j >= n) continue;
            if (ti[k] != tj[k]) {
                atomicAdd(&cb[(int64_t)i * n + j], acc[k][e]);
                atomicAdd(&cb[(int64_t)j * n + i], acc[k][e]);
            } else {
                atomicAdd(&cb[(int64_t)i * n + j], acc[k][e]);   // diagonal tiles hold both triangles themselves
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Batched symmetric eigensolver: parallel cyclic Jacobi (round-robin ordering), one CTA per matrix.
// Rotations are skipped when |a_pq| <= tol sqrt(|a_pp a_qq|) (relative criterion => small eigenvalues of positive
// definite matrices are found to high relative accuracy); the sweep loop stops when a full sweep applies no rotation.
//
// A step rotates N / 2 disjoint index pairs I_k = (p_k, q_k) at once: A' = J^T A J, V' = V J.  Seen in 2 x 2 blocks,
// A'[I_k, I_l] = J_k^T A[I_k, I_l] J_l, so the two-sided update is done in ONE pass with one thread per block: 4 loads,
// both rotations in registers, and -- A being symmetric -- only the blocks k <= l are computed and only the upper
// triangle of A is kept (element (i, j) lives at [min(i, j)][max(i, j)]).  The eigenvector matrix is kept TRANSPOSED (row e = vector e), so
// its column rotations are 16-byte vector operations on two contiguous rows.  Against the previous form (a column pass
// over A and V with strided scalar accesses, a barrier, a row pass over A, a barrier) this issues ~2.3x fewer
// instructions per step and needs two barriers instead of three; the kernel is issue bound.
// ------------------------------------------------------------------------------------------------
constexpr int kJacThreads = 256;
#ifdef PMD_TUNE
__device__ unsigned long long g_jac_sweep_hist[2][64];   // development builds: histogram of sweeps used ([float32 sweeps?][count])
#endif
// 2 x 2 blocks per thread: (N/2)(N/2 + 1)/2 <= 528 for n <= 64 (3 items), <= 1596 for n <= 112 (7 items)

template <typename S> struct JacVec;
template <> struct JacVec<float> { using type = float4; static constexpr int W = 4; };
template <> struct JacVec<double> { using type = double2; static constexpr int W = 2; };

__device__ __forceinline__ void jac_rot_rows(float4& x, float4& y, float c, float s) {
    const float4 a = x, b = y;
    x = make_float4(c * a.x - s * b.x, c * a.y - s * b.y, c * a.z - s * b.z, c * a.w - s * b.w);
    y = make_float4(s * a.x + c * b.x, s * a.y + c * b.y, s * a.z + c * b.z, s * a.w + c * b.w);
}
__device__ __forceinline__ void jac_rot_rows(double2& x, double2& y, double c, double s) {
    const double2 a = x, b = y;
    x = make_double2(c * a.x - s * b.x, c * a.y - s * b.y);
    y = make_double2(s * a.x + c * b.x, s * a.y + c * b.y);
}

// S = working precision of the sweeps: double (default) or float (sketch-stage subspaces).
template <typename S, int kJacMaxItems>
__global__ void __launch_bounds__(kJacThreads, kJacMaxItems <= 3 ? 5 : 1)
jacobi_eigh_kernel(double* __restrict__ cmat, int n, int mode, int max_sweeps, double* __restrict__ w_out,
                   float* __restrict__ vec_out) {
    extern __shared__ __align__(16) unsigned char jsm_raw[];
    using VT = typename JacVec<S>::type;
    constexpr int VW = JacVec<S>::W;
    // rotation threshold (relative to sqrt(a_pp a_qq)) and the early-exit bound: a sweep whose LARGEST rotated element was
    // below kQuad leaves off-diagonal elements of second order (<= kQuad^2 < kTol), so the sweep that would only verify
    // convergence is not run.  The eigenvectors are returned in float32 (6e-8), the float32 sweeps feed the sketch stage.
    const S kTol = sizeof(S) == 8 ? (S)1e-14 : (S)1e-6;
    const S kQuad = sizeof(S) == 8 ? (S)3e-8 : (S)3e-4;
    const int N = n + (n & 1);           // even size: index n (if any) is a padding row / column of zeros
    const int ld = N | 1;                // odd pitch of A
    const int ldv = (N + VW - 1) / VW * VW;
    const int half = N / 2;
    const size_t a_bytes = ((size_t)N * ld * sizeof(S) + 15) / 16 * 16;
    S* A = reinterpret_cast<S*>(jsm_raw);                       // [N][ld]
    S* Vt = reinterpret_cast<S*>(jsm_raw + a_bytes);            // [N][ldv], row e = eigenvector e
    S* cc = Vt + (size_t)N * ldv;                               // [half]
    S* ss = cc + half;                                          // [half]
    int* pq = reinterpret_cast<int*>(ss + half);                // [half]: p | q << 8 | rotated << 16
    __shared__ int n_rot, n_big;
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.x;
    double* cb = cmat + b * (int64_t)n * n;

    for (int idx = tid; idx < N * ld; idx += kJacThreads) {
        const int i = idx / ld, j = idx - i * ld;
        A[idx] = (i < n && j < n) ? (S)cb[i * n + j] : (S)0;
    }
    for (int idx = tid; idx < N * ldv; idx += kJacThreads) {
        const int i = idx / ldv, j = idx - i * ldv;
        Vt[idx] = (i == j) ? (S)1 : (S)0;
    }
    // the (k, l), k <= l, blocks of this thread: the upper triangle of the half x half block grid folded into a rectangle
    const int fold_cols = (half & 1) ? half : half + 1;
    const int n_blocks = half * (half + 1) / 2;
    int kl[kJacMaxItems];
#pragma unroll
    for (int it = 0; it < kJacMaxItems; ++it) {
        const int m = tid + it * kJacThreads;
        kl[it] = -1;
        if (m < n_blocks) {
            const int a = m / fold_cols, c = m - a * fold_cols;
            int k, l;
            if (a + c < half) { k = a; l = a + c; }
            else if (half & 1) { k = half - a; l = c; }
            else { k = half - 1 - a; l = k + (c - (half - a)); }
            kl[it] = k | (l << 8);
        }
    }
    const int v_chunks = ldv / VW, v_items = half * v_chunks;
    __syncthreads();

    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        if (tid == 0) n_rot = n_big = 0;
        __syncthreads();
        for (int step = 0; step < N - 1; ++step) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = N - 1; q = step; }
                else { p = (step + tid) % (N - 1); q = (step - tid + (N - 1)) % (N - 1); }
                if (p > q) { const int tmp = p; p = q; q = tmp; }
                S c = 1, s = 0;
                int rot = 0;
                if (q < n) {
                    const S apq = A[p * ld + q], app = A[p * ld + p], aqq = A[q * ld + q];
                    const S scale = sqrt(fabs(app * aqq));
                    if (apq != (S)0 && fabs(apq) > kTol * scale) {
                        if (fabs(apq) > kQuad * scale) n_big = 1;     // (benign race: every writer stores 1)
                        const S tau = (aqq - app) / ((S)2 * apq);
                        const S t = (tau >= (S)0 ? (S)1 : (S)-1) / (fabs(tau) + sqrt((S)1 + tau * tau));
                        c = (S)1 / sqrt((S)1 + t * t);
                        s = t * c;
                        rot = 1;
                    }
                }
                pq[tid] = p | (q << 8) | (rot << 16);
                cc[tid] = c;
                ss[tid] = s;
                if (rot) atomicAdd(&n_rot, 1);
            }
            __syncthreads();
            // A' = J^T A J, one 2 x 2 block per item
#pragma unroll
            for (int it = 0; it < kJacMaxItems; ++it) {
                if (kl[it] < 0) break;
                const int k = kl[it] & 255, l = kl[it] >> 8;
                const int pqk = pq[k], pql = pq[l];
                if (((pqk | pql) >> 16) == 0) continue;           // neither pair rotates
                const int pk = pqk & 255, qk = (pqk >> 8) & 255, pl = pql & 255, ql = (pql >> 8) & 255;
                const S ck = cc[k], sk = ss[k], cl = cc[l], sl = ss[l];
                // only the upper triangle of A is live: element (i, j) is kept at [min(i, j)][max(i, j)] (half the stores of
                // a mirrored update; the kernel is bound by shared-memory bandwidth)
                const int i00 = min(pk, pl) * ld + max(pk, pl), i01 = min(pk, ql) * ld + max(pk, ql);
                const int i10 = min(qk, pl) * ld + max(qk, pl), i11 = min(qk, ql) * ld + max(qk, ql);
                const S a00 = A[i00], a01 = A[i01], a10 = A[i10], a11 = A[i11];   // k == l: i10 == i01
                // columns:  t_i0 = cl a_i0 - sl a_i1,  t_i1 = sl a_i0 + cl a_i1
                const S t00 = cl * a00 - sl * a01, t01 = sl * a00 + cl * a01;
                const S t10 = cl * a10 - sl * a11, t11 = sl * a10 + cl * a11;
                // rows:     b_0j = ck t_0j - sk t_1j,  b_1j = sk t_0j + ck t_1j
                const S b00 = ck * t00 - sk * t10, b01 = ck * t01 - sk * t11;
                const S b10 = sk * t00 + ck * t10, b11 = sk * t01 + ck * t11;
                A[i00] = b00;
                A[i11] = b11;
                if (k == l) {
                    A[i01] = (S)0;                                 // the annihilated element (p_k < q_k: i01 is its upper copy)
                } else {
                    A[i01] = b01;
                    A[i10] = b10;
                }
            }
            // V' = V J on the transposed copy: rows p_k, q_k
            for (int m = tid; m < v_items; m += kJacThreads) {
                const int k = m / v_chunks, ch = m - k * v_chunks;
                const int pqk = pq[k];
                if ((pqk >> 16) == 0) continue;
                const int pk = pqk & 255, qk = (pqk >> 8) & 255;
                VT* rp = reinterpret_cast<VT*>(Vt + pk * ldv) + ch;
                VT* rq = reinterpret_cast<VT*>(Vt + qk * ldv) + ch;
                VT x = *rp, y = *rq;
                jac_rot_rows(x, y, cc[k], ss[k]);
                *rp = x;
                *rq = y;
            }
            __syncthreads();
        }
        const int rots = n_big ? n_rot : 0;
        __syncthreads();
#ifdef PMD_TUNE
        if (tid == 0 && (rots == 0 || sweep == max_sweeps - 1)) atomicAdd(&g_jac_sweep_hist[sizeof(S) == 4][min(sweep + 1, 63)], 1ull);
#endif
        if (rots == 0) break;
    }

    // sort descending (rank by counting), write eigenvalues and (scaled) eigenvectors
    double wmax = -1e300;
    for (int i = 0; i < n; ++i) wmax = fmax(wmax, (double)A[i * ld + i]);
    for (int idx = tid; idx < n * n; idx += kJacThreads) {
        const int i = idx / n, r = idx % n;  // element r of eigenvector i
        const double wi = (double)A[i * ld + i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const double wj = (double)A[j * ld + j];
            rank += (wj > wi) || (wj == wi && j < i);
        }
        double scale = 1.0;
        if (mode == 1) scale = (wi > wmax * 1e-24 && wi > 0.0) ? rsqrt(wi) : 0.0;
        vec_out[b * (int64_t)n * n + (int64_t)r * n + rank] = (float)((double)Vt[i * ldv + r] * scale);
        if (r == 0) w_out[b * (int64_t)n + rank] = wi;
    }
}

}  // namespace pmd

extern "C" int pmd_gram_f64(const float* a, int64_t batch, int64_t n, int64_t m_len, int64_t batch_stride,
                            int64_t row_stride, int64_t inner_stride, double* c, void* stream) {
    const char* fn = "pmd_gram_f64";
    PMD_REQUIRE(a && c, fn, "null pointer");
    PMD_REQUIRE(batch > 0 && batch <= 65535 && n > 0 && n <= 112 && m_len > 0, fn, "bad size (n <= 112, batch <= 65535)");
    PMD_REQUIRE(inner_stride == 1 || row_stride == 1, fn, "one of row_stride / inner_stride must be 1");
    // enough CTAs to fill the machine, at least 4 inner tiles each
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>((m_len + 4 * pmd::kGramMT - 1) / (4 * pmd::kGramMT),
                                                            (4 * 148 + batch - 1) / batch));
    int64_t m_per = (m_len + splits - 1) / splits;
    m_per = (m_per + pmd::kGramMT - 1) / pmd::kGramMT * pmd::kGramMT;
    splits = (m_len + m_per - 1) / m_per;
    if (inner_stride == 1 && (row_stride % 4) == 0) {
        // contiguous rows: FP64 tensor cores
        const int nt8 = (int)(n + 7) / 8;
        const size_t smem_d = (size_t)8 * nt8 * pmd::kDGLd * sizeof(float);
        auto kern = nt8 * (nt8 + 1) / 2 <= 32 ? pmd::gram_dmma_kernel<4> : pmd::gram_dmma_kernel<14>;
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
        if (e2 != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e2)); return (int)e2; }
        dim3 grid_d((unsigned)splits, (unsigned)batch);
        kern<<<grid_d, pmd::kDGThreads, smem_d, (cudaStream_t)stream>>>(a, (int)n, m_len, batch_stride, row_stride, m_per, c);
        return pmd::check_launch(fn);
    }
    const int nt = (int)(n + 1) / 2;
    const size_t smem = (size_t)2 * nt * (pmd::kGramMT + 1) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(pmd::gram_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    dim3 grid((unsigned)splits, (unsigned)batch);
    pmd::gram_f64_kernel<<<grid, pmd::kGramThreads, smem, (cudaStream_t)stream>>>(a, (int)n, m_len, batch_stride, row_stride,
                                                                                 inner_stride, m_per, c);
    return pmd::check_launch(fn);
}

#ifdef PMD_TUNE
extern "C" int pmd_debug_jacobi_hist(unsigned long long* out, int reset) {   // development builds only (not in the header)
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, pmd::g_jac_sweep_hist, sizeof(unsigned long long) * 128);
    if (reset) {
        unsigned long long z[128] = {0};
        cudaMemcpyToSymbol(pmd::g_jac_sweep_hist, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" int pmd_jacobi_eigh(double* c, int64_t batch, int64_t n, int mode, int sweeps_f32, double* w, float* vecs,
                               void* stream) {
    const char* fn = "pmd_jacobi_eigh";
    PMD_REQUIRE(c && w && vecs, fn, "null pointer");
    PMD_REQUIRE(batch > 0 && n > 0 && n <= 112, fn, "bad size (n <= 112)");
    PMD_REQUIRE(mode == 0 || mode == 1, fn, "mode must be 0 or 1");
    const int N = (int)n + ((int)n & 1), ld = N | 1, half = N / 2;
    const size_t es = sweeps_f32 ? sizeof(float) : sizeof(double);
    const int vw = (int)(16 / es), ldv = (N + vw - 1) / vw * vw;
    const size_t smem = ((size_t)N * ld * es + 15) / 16 * 16 + (size_t)N * ldv * es + (size_t)2 * half * es + (size_t)half * sizeof(int);
    cudaStream_t st = (cudaStream_t)stream;
    void (*k)(double*, int, int, int, double*, float*);
    if (sweeps_f32)
        k = n <= 64 ? pmd::jacobi_eigh_kernel<float, 3> : pmd::jacobi_eigh_kernel<float, 7>;
    else
        k = n <= 64 ? pmd::jacobi_eigh_kernel<double, 3> : pmd::jacobi_eigh_kernel<double, 7>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pmd::set_error(std::string(fn) + ": " + cudaGetErrorString(e)); return (int)e; }
    k<<<(unsigned)batch, pmd::kJacThreads, smem, st>>>(c, (int)n, mode, 40, w, vecs);
    return pmd::check_launch(fn);
}
