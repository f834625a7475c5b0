/*
 * pmd_sm100.h -- C ABI of libpmd_sm100.so: the B200 (sm_100a) kernels behind the hot path of
 * apasarkar/localmd (penalized / local matrix decomposition of imaging movies).
 *
 * The reference has no FFI of its own: its boundary is the Python API (localmd_decomposition,
 * PMDArray).  Each entry point below replaces one jitted/XLA computation or host loop of the
 * reference; the `replaces:` line names it as file:line under the reference checkout.  The Python
 * host (localmd_b200/) binds these with ctypes and mirrors the reference's Python interface.
 *
 * Conventions (all entry points):
 *   - every array argument is a DEVICE pointer owned by the caller (torch tensors); the library
 *     never allocates, frees or keeps device memory and holds no global state;
 *   - sizes are int64_t; `stream` is a cudaStream_t passed as void*; work is enqueued
 *     asynchronously on that stream; the call is re-entrant and thread safe;
 *   - return value: 0 ok, <0 invalid argument, >0 a cudaError_t; pmd_last_error() returns a
 *     thread-local message for the last non-zero return on the calling thread;
 *   - movies are frame-major: element (frame f, pixel p) at base[f*d + p], p = row*d2 + col
 *     (the physical layout of the reference's (T,d1,d2) dataset); `dtype` selects the element type;
 *   - "block" = one overlapping spatial tile (decomposition.py:723-739); block-local pixel index is
 *     q = qi*bw + qj, global pixel (i0+qi)*d2 + j0+qj.
 */
#ifndef PMD_SM100_H
#define PMD_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* movie element types accepted wherever a `dtype` argument appears */
enum pmd_dtype { PMD_F32 = 0, PMD_U16 = 1, PMD_I16 = 2, PMD_U8 = 3, PMD_F64 = 4, PMD_I32 = 5 };

const char* pmd_last_error(void);
int pmd_abi_version(void);

/* K1  per-pixel mean and Welch high-band noise estimate, one streaming pass.
 * replaces: pmd_loader.py:203-291 (PMDLoader._calculate_mean_and_normalizer) and
 *           preprocessing_utils.py:10-40 (get_mean_and_noise, get_mean_chunk, get_noise_estimate).
 * movie: t_local frames of d pixels.  Frames are cut into chunks of 1024 starting at frame 0 of
 * this buffer (n_chunks = ceil(t_local/1024)).  Outputs, both [n_chunks][d] float32:
 *   mean_part[c][p]  = (sum of the chunk's frames at p) / t_total
 *   noise_part[c][p] = Welch estimate of chunk c (0 where the chunk has < 256 frames)
 * tab: 772 float32 built by the host (localmd_b200/_tables.py: welch_fft_tables): the periodic Hann window
 * [256], the FFT twiddles (cos, -sin)(2 pi j/128) [128][2] and the split twiddles (cos, sin)(2 pi k/256)
 * [130][2] (entries 129 unused). */
int pmd_stats_pass(const void* movie, int dtype, int64_t t_local, int64_t d, int64_t t_total, const float* tab,
                   float* mean_part, float* noise_part, void* stream);

/* K1 on the tensor cores (csrc/stats_tc.cu): same contract and outputs as pmd_stats_pass, for movies whose frame pitch
 * d * sizeof(element) is a multiple of 16 bytes and whose base address is 16-byte aligned (2-D TMA boxes of 16 frames x
 * 128 pixels).  The segment transform runs as TF32 + bf16-pair tcgen05 MMAs on the radix-2 decimation-in-frequency form
 * (even bins from w x[t] + (1 - w) x[t + 128], odd bins from w x[t] - (1 - w) x[t + 128]) against two constant (64 x 128)
 * matrices.  tab: 131584 bytes built by the host (localmd_b200/_tables.py: welch_tc_tables): the matrices in the
 * shared-memory operand layout (TF32 parts, then bf16 pair parts; K-major SWIZZLE_128B) and the first half of the
 * periodic Hann window as float32.
 * replaces: pmd_loader.py:203-291 and preprocessing_utils.py:10-40 (as pmd_stats_pass). */
int pmd_stats_pass_tc(const void* movie, int dtype, int64_t t_local, int64_t d, int64_t t_total, const void* tab,
                      float* mean_part, float* noise_part, void* stream);

/* gather + standardise frames: out[i][p] = (movie[frames[i]][p] - mean[p]) / stdv[p]  (float32).
 * replaces: pmd_loader.py:293-298 (temporal_crop_standardized) and the first two lines of
 *           standardize_and_filter, pmd_loader.py:374-377. */
int pmd_standardize_frames(const void* movie, int dtype, int64_t d, const int64_t* frames, int64_t n_frames,
                           const float* mean, const float* stdv, float* out, void* stream);

/* batched Gram matrices in float64:  C[b] = A[b] A[b]^T  (or A^T A), fp32 inputs, fp64 accumulate.
 * A[b] element (i, m) at a + b*batch_stride + i*row_stride + m*inner_stride, i < n (n <= 112),
 * m < m_len.  C is [batch][n][n] double and must be zeroed by the caller (partial sums are added
 * atomically).  Rows that are contiguous in memory (inner_stride 1, row_stride a multiple of 4) run on the FP64
 * tensor cores (mma.sync m8n8k4); other layouts on FMA.  Building block of every small orthogonalisation / SVD below.
 * replaces: the normal-equation half of jnp.linalg.qr / jnp.linalg.svd calls at
 *           decomposition.py:64,66,301,315,319 and pmd_loader.py:58,60. */
int pmd_gram_f64(const float* a, int64_t batch, int64_t n, int64_t m_len, int64_t batch_stride,
                 int64_t row_stride, int64_t inner_stride, double* c, void* stream);

/* batched symmetric eigensolver (cyclic parallel Jacobi in shared memory, one CTA per matrix, n <= 112; a step's disjoint
 * rotations are applied as one pass over the 2 x 2 blocks of the upper triangle, the eigenvectors are kept transposed).
 * Rotations are skipped below 1e-14 (float32 sweeps: 1e-6) of sqrt(a_pp a_qq); the sweeps stop after a sweep whose largest
 * rotated element was below the second-order bound 3e-8 (3e-4) or that rotated nothing.
 * c: [batch][n][n] double (destroyed).  Outputs: w [batch][n] double eigenvalues, descending;
 * vecs [batch][n][n] float32, column j = eigenvector j, scaled per `mode`:
 *   0: orthonormal eigenvectors E
 *   1: E * diag(1/sqrt(w))  ("whitening": X*vecs has orthonormal columns when c = X^T X);
 *      columns with w_j <= w_0 * 1e-24 are zeroed
 * sweeps_f32 != 0: the rotations run in float32 (reference-level accuracy, 1e-7 of the largest eigenvalue) -- used
 * for the sketch-stage SVD of decomposition.py:66, whose result only seeds the temporal basis; the float64 CUDA-core
 * rate of this GPU is ~1/64 of float32.
 * replaces: the small LAPACK SVD/eigh factorisations inside decomposition.py:66,301,315,319 and
 *           pmd_loader.py:60. */
int pmd_jacobi_eigh(double* c, int64_t batch, int64_t n, int mode, int sweeps_f32, double* w, float* vecs,
                    void* stream);

/* gather + standardise + TRANSPOSE frames into the pixel-major init movie used by the block stage:
 *   out[p*ld + i] = (movie[frames[i]][p] - mean[p]) / stdv[p]   (i < n_frames; columns n_frames..ld-1 are zeroed)
 * replaces: pmd_loader.py:348-371 (temporal_crop_with_filter's frame loads) + 374-377.  ld >= n_frames, a
 * multiple of 4 (16-byte rows). */
int pmd_standardize_frames_t(const void* movie, int dtype, int64_t d, const int64_t* frames, int64_t n_frames,
                             const float* mean, const float* stdv, float* out, int64_t ld, void* stream);

/* batched in-place orthonormalisation of the first n columns of x [batch][m][ldx] (float32), one CTA per matrix held
 * in shared memory (n <= 64; Gram and triangular solve on the FP64 tensor cores, mma.sync m8n8k4):  if g_ext != NULL first  X <- X L_g^-T  with g_ext[b] = L_g L_g^T ([batch][n][n]
 * float64: used with the Gram of the temporal components so that X = block * V^T becomes block * (orthonormal
 * temporal basis)^T, decomposition.py:301-306); then `passes` rounds of CholQR (float64 Gram, in-kernel Cholesky,
 * triangular solve).  Numerically dependent columns are set to zero.
 * replaces: jnp.linalg.qr at decomposition.py:64 and the basis-producing SVDs at 301 and 315. */
int pmd_block_orth(float* x, int64_t batch, int64_t m, int64_t n, int64_t ldx, const double* g_ext,
                   int64_t passes, void* stream);

/* Background removal on the pixel-major init movie yt [d][ld] (ld a multiple of 4, padding columns included):
 *   pmd_bg_project_t: part[g][c][f] = sum over the pixels of range g of bg[c][p] * yt[p][f]   (n_ranges equal pixel
 *                     ranges; the caller sums the partials over g in a fixed order -> vbg [k][ld], deterministic)
 *   pmd_bg_remove_t:  yt[p][f] -= sum_c bg[c][p] * vbg[c][f]
 * bg: [k][d] float32 orthonormal background rows, 1 <= k <= 16 (pmd_bg_project_t alone also takes 17 <= k <= 32: the
 *     sketch coefficients of the background rSVD below).
 * replaces: pmd_loader.py:386-387 (standardize_and_filter: temporal projection onto the spatial background basis
 *           and its subtraction) on the init frames. */
int pmd_bg_project_t(const float* yt, int64_t ld, int64_t d, const float* bg, int64_t k, int64_t n_ranges, float* part,
                     void* stream);
int pmd_bg_remove_t(float* yt, int64_t ld, int64_t d, const float* bg, int64_t k, const float* vbg, void* stream);

/* Background basis: the two skinny contractions of the randomised SVD of the sampled standardised frames
 * (pmd_loader.py:46-68, called from 300-314) on their pixel-major copy yt [d][ld] (pmd_standardize_frames_t):
 *   pmd_rows_sketch:       y[p][j] = sum_{f < n} yt[p][f] * omega[f][j]        omega [n][l] row-major, 1 <= l <= 32
 *                          (the coefficient pass  q^T yt  is pmd_bg_project_t with k = l)
 *   pmd_rows_times_small:  out[b][p][c] = sum_{j < k} x[b][p][j] * m[b][j][c]   k, nc <= 32; transposed != 0 writes
 *                          out[b][c][p] instead (row pitch ldo either way).  Used for the orthonormalisation passes
 *                          (x @ whitening transform) and the rotation into singular vectors.  Not in place.
 *   pmd_chol_whiten:       t[b] = L^-T (float32, n x n) of the Cholesky factor g[b] = L L^T of a float64 Gram matrix,
 *                          n <= 32: x @ t is an orthonormal basis of the range of x when g = x^T x (CholQR; two rounds
 *                          replace the QR of pmd_loader.py:59).  Numerically dependent columns give zero columns.
 * replaces: the jnp.matmul / jnp.linalg.qr / svd chain of pmd_loader.py:55-68 (random_svd of the background frames). */
int pmd_chol_whiten(const double* g, int64_t batch, int64_t n, float* t, void* stream);
int pmd_rows_sketch(const float* yt, int64_t ld, int64_t d, int64_t n, const float* omega, int64_t l, float* y, int64_t ldy,
                    void* stream);
int pmd_rows_times_small(const float* x, int64_t ldx, int64_t d, int64_t k, const float* m, int64_t ldm, int64_t nc, float* out,
                         int64_t ldo, int transposed, int64_t batch, int64_t batch_stride_x, int64_t batch_stride_m,
                         int64_t batch_stride_o, void* stream);

/* 2x2(-ish) average pooling + temporal averaging of every block of the standardised init movie.
 * replaces: decomposition.py:192-232 (downsample_average_pooling) + 283-290.
 * yt: pixel-major init movie [d][ld] float32 (frame f of pixel p at yt[p*ld+f]), t frames used.
 * starts: [nb][2] int32 (i0, j0).  Output bta [nb][P][t/taf] float32 with
 * P = ceil(bh/saf)*ceil(bw/saf), pooled pixel index pi*ceil(bw/saf)+pj, XLA 'SAME' padding. */
int pmd_block_pool_tavg(const float* yt, int64_t ld, int64_t t, int64_t d2, const int32_t* starts, int64_t nb,
                        int64_t bh, int64_t bw, int64_t saf, int64_t taf, float* bta, void* stream);

/* same pooling, but ALSO keeps the pooled block at full time resolution:  pooled [nb][P][ld] (columns t..ld-1 zero)
 * = B_ds of decomposition.py:279, so that U_ds^T B_ds (295-298) contracts over P pooled pixels (pmd_block_project
 * with movie_batch_stride = P*ld, d2 = ceil(bw/saf), starts = 0).  t a multiple of taf. */
int pmd_block_pool_full(const float* yt, int64_t ld, int64_t t, int64_t d2, const int32_t* starts, int64_t nb,
                        int64_t bh, int64_t bw, int64_t saf, int64_t taf, float* pooled, float* bta, void* stream);

/* spread a pooled spatial basis back to full resolution: w[b][q][c] = uds[b][pool(q)][c] / count(pool(q)),
 * so that w^T * block == uds^T * pooled(block)  (decomposition.py:295-298 without materialising the
 * pooled block).  uds: [nb][P][r], w: [nb][bh*bw][rp] (rp >= r, multiple of 4, padding zeroed). */
int pmd_block_unpool(const float* uds, int64_t nb, int64_t bh, int64_t bw, int64_t saf, int64_t r, int64_t rp,
                     float* w, void* stream);

/* streaming block projection  out[b][c][f] = sum_q w[b][q][c] * Y_b[q][f]   (c < r, f < ldo).
 * replaces: decomposition.py:295-298 (u^T * pooled block), 318 (u_final^T * block), 390-407.
 * movie_t: pixel-major float32 [d][ld] (+ b*movie_batch_stride elements for block b: 0 = all blocks share one
 * movie; non-zero is used by the threshold simulation where every "block" is its own tiny movie).  ld is a
 * multiple of 4 and columns t..ld-1 hold zeros.  w: [nb][bh*bw][rp]; out: [nb][r][ldo], ldo <= ld. */
int pmd_block_project(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                      const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w, int64_t r,
                      int64_t rp, float* out, int64_t ldo, void* stream);

/* the same block projection on the tcgen05 tensor cores (3xTF32, float32-class accuracy, accumulators in tensor
 * memory).  w_hi / w_lo: [nb][bh*bw][rp] float32 with w = w_hi + w_lo and w_hi exactly representable in TF32 (low 13
 * mantissa bits zero); rp <= 64.  Same outputs as pmd_block_project. */
int pmd_block_project_tc(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                         const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w_hi,
                         const float* w_lo, int64_t r, int64_t rp, float* out, int64_t ldo, void* stream);

/* the block projection with the movie operand in tensor memory (csrc/blocks_ts.cu): raw [32 pixels x 128 frames] tiles
 * by 2-D TMA boxes, split in registers and written to tensor memory (tcgen05.st), TF32 + bf16-pair MMAs (float32-class
 * accuracy), persistent CTAs with double-buffered accumulators.  w: [nb][bh*bw][rp] float32 (unsplit; rp <= 64);
 * workspace: pmd_block_project_ts_workspace_bytes(nb, bh, bw) bytes of device memory for the packed coefficient
 * images (16-byte aligned); n_rows: number of pixel rows of movie_t (all batches); movie_batch_stride must be a multiple
 * of ld; bw even.  Same outputs as pmd_block_project. */
int64_t pmd_block_project_ts_workspace_bytes(int64_t nb, int64_t bh, int64_t bw);
int pmd_block_project_ts(const float* movie_t, int64_t movie_batch_stride, int64_t n_rows, int64_t ld, int64_t d2,
                         const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* w, int64_t r, int64_t rp,
                         void* workspace, float* out, int64_t ldo, void* stream);

/* streaming spatial projection  s[b][q][c] = sum_f Y_b[q][f] * v[b][c][f]   (all ldv frames of v).
 * replaces: decomposition.py:304-306 (block * v_basis^T).   v: [nb][r][ldv] (ldv multiple of 4, padding
 * zero); s: [nb][bh*bw][rp], rp <= 64. */
int pmd_block_spatial(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                      const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* v, int64_t ldv,
                      int64_t r, int64_t rp, float* s, void* stream);

/* the same spatial projection on the tcgen05 tensor cores (3xTF32, K-major SWIZZLE_128B operands, accumulators in
 * tensor memory over the whole frame loop).  Same arguments and output as pmd_block_spatial. */
int pmd_block_spatial_tc(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                         const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* v, int64_t ldv,
                         int64_t r, int64_t rp, float* s, void* stream);

/* the spatial projection with the movie operand in tensor memory (csrc/blocks_ts.cu): every thread streams its pixel's
 * frames straight into registers, splits them into TF32 hi + bf16 pairs and writes the operand into tensor memory; the
 * temporal rows are converted once per 32-frame chunk into the shared-memory operand image; accumulators of up to four
 * 128-pixel tiles stay in tensor memory over the whole frame loop; persistent CTAs.  Same arguments and output as
 * pmd_block_spatial (any block size; nb unlimited). */
int pmd_block_spatial_ts(const float* movie_t, int64_t movie_batch_stride, int64_t ld, int64_t d2,
                         const int32_t* starts, int64_t nb, int64_t bh, int64_t bw, const float* v, int64_t ldv,
                         int64_t r, int64_t rp, float* s, void* stream);

/* roughness statistics of every component of every block + the keep-through-first-failure rule.
 * replaces: evaluation.py:84-126 (spatial/temporal_roughness_stat), 133-192, 195-222
 *           (filter_by_failures) and decomposition.py:502-506.
 * u: [nb][bh*bw][rp] spatial components, v: [nb][r][ldv] temporal components (first t columns used).
 * Outputs: sstat, tstat [nb][r] float32; ranks [nb] int32 (number of kept leading components). */
int pmd_block_stats_rank(const float* u, const float* v, int64_t nb, int64_t bh, int64_t bw, int64_t r,
                         int64_t rp, int64_t t, int64_t ldv, float thr_s, float thr_t, int64_t max_fail,
                         float* sstat, float* tstat, int32_t* ranks, void* stream);

/* weighted assembly of the sparse spatial matrix in block-component form.
 * replaces: decomposition.py:811-853 (pyramid weighting, COO construction, division by the summed
 *           weights).  For block b, kept component c (< ranks[b]) and block pixel q:
 *   val = (1.0/cumw[pix]) * ((double)u[b][q][c] * (double)bw_img[q])        (float64, as the reference)
 * written to uvals64[(col0[b]+c)*bpix + q]  and, as float32, to uvals32[...]. */
int pmd_assemble_u(const float* u, int64_t nb, int64_t bh, int64_t bw, int64_t rp, const int32_t* starts,
                   const int32_t* ranks, const int64_t* col0, const float* block_weights, const double* cumw,
                   int64_t d2, double* uvals64, float* uvals32, void* stream);

/* K7a  full-movie projection onto the local (block-supported) columns of U:
 *   z[col0[b]+c][f] = sum_q uvals32[(col0[b]+c)*bpix+q] * (movie[f][pix(b,q)] - mean[pix]) * inv_std[pix]
 * replaces: pmd_loader.py:316-346, 392-414 (v_projection: reshape, (Y-mu)/sigma, BCOO U^T @) for
 *           the block columns.  mean/inv_std may be NULL (no standardisation; float32 movies only:
 *           used for U^T (U M) in the whitening step, decomposition.py:974-981).
 * tasks: [n_tasks][2] int32 = (block, first component) for every group of <= 4 kept components
 * (built by the host from ranks).  z: [n_cols_total][ldz] float32, written for f < t; must be zeroed by
 * the caller when bh*bw > 512 (pixel slabs are then accumulated atomically). */
int pmd_project_local(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* starts,
                      int64_t nb, int64_t bh, int64_t bw, const int32_t* ranks, const int64_t* col0,
                      const int32_t* tasks, int64_t n_tasks, const float* uvals32, const float* mean,
                      const float* inv_std, float* z, int64_t ldz, void* stream);

/* K7a (v2)  same contraction as pmd_project_local, organised by SUPERTILES of neighbouring blocks whose
 * pixel union is staged once per frame sub-tile in shared memory (centred and scaled there).
 * tiles: [n_tiles][4] int32 = (r0, c0, rh, rw) pixel region of each supertile (rh*rw <= 2048);
 * task_ptr: [n_tiles+1] int32 offsets into tasks; tasks: [n_tasks][4] int32 = (row offset of the block inside
 * the region, column offset, first output column, number of components 1..4).  Built by the host
 * (localmd_b200/ops.py: make_supertiles).  bh*bw <= 512.  Every z element of a listed column is written once. */
int pmd_project_supertile(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* tiles,
                          int64_t n_tiles, const int32_t* task_ptr, const int32_t* tasks, int64_t bh, int64_t bw,
                          int64_t max_region_h, int64_t max_region_w, const float* uvals32, const float* mean,
                          const float* inv_std, float* z, int64_t ldz, void* stream);

/* K7 (v3)  the whole projection  z = U^T ((Y - mean) * inv_std)  in ONE streaming pass over the movie: local
 * (block-supported) and dense background columns together.  A CTA walks a column strip of the field of view
 * (all rows, <= 48 pixels wide) for 256 frames, one pixel row at a time; its 8 warps own "tasks" packed into
 * slots by the host (localmd_b200/ops.py: make_strips):
 *   items:    [n_items][8] int32  = (c0, rw, first slot_ptr entry, n_rows, bg partial index, first row, 0, 0)
 *   slot_ptr: 9 consecutive int32 per item: tasks of warp w are slot_ptr[first+w] .. slot_ptr[first+w+1]-1
 *   tasks:    [n_tasks][12] int32 = (first row, column offset in the strip, rows, width, output column,
 *             number of components 1..8, padded components 4|8, floats per row of its U pack, U pack offset
 *             (low, high 32 bits), kind 0 local | 1 background, 0), row ranges of one slot disjoint, ascending
 *   upack:    float32 U values per task as [row][pixel][padded component] (zero padded)
 * Local tasks write z[col..col+nc) (every element once); background tasks write the partial sums of their
 * strip to zbg[bg partial][component][frame] (the caller adds the partials).  mean / inv_std may be NULL.
 * replaces: pmd_loader.py:316-346, 392-414 (v_projection / v_projection_routine) in full. */
int pmd_project_stream(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* items,
                       int64_t n_items, const int32_t* slot_ptr, const int32_t* tasks, int64_t max_rw,
                       const float* upack, const float* mean, const float* inv_std, float* z, int64_t ldz,
                       float* zbg, int64_t ldzbg, int64_t bg_stride, void* stream);

/* HOST function (every pointer is a HOST pointer, no device work): builds the items / slot_ptr / tasks tables of
 * pmd_project_stream from the block grid (row_starts x col_starts, blocks numbered row-major), the kept ranks and
 * first output columns of the blocks and the number of dense background columns.  g_fixed > 0 forces that many block
 * columns per strip, otherwise the strip width minimising the streamed pixels is chosen.  Outputs are caller
 * allocated: items [cap_items][8], slot_ptr [cap_items*9], tasks [cap_tasks][12], local8 / local4 [cap_tasks][2]
 * int64 (first column, n comps) in the order their U values must be packed, counts[8] = (n_items, n_slot_ptr, n_tasks,
 * n_local8, n_local4, n_parts, max strip width, floats of the U pack); counts[0] == 0: geometry not supported.
 * replaces: nothing in the reference (host bookkeeping of the new projection kernel). */
int pmd_make_strips(const int32_t* row_starts, int64_t nbr, const int32_t* col_starts, int64_t nbc, int64_t bh,
                    int64_t bw, int64_t d1, int64_t d2, const int64_t* ranks, const int64_t* col0, int64_t n_bg,
                    int64_t g_fixed, int32_t* items_out, int64_t cap_items, int32_t* slot_ptr_out,
                    int32_t* tasks_out, int64_t cap_tasks, int64_t* local8_out, int64_t* local4_out,
                    int64_t* counts);

/* K7 on the tensor cores (tcgen05, sm_100a): the same projection as pmd_project_stream, computed as
 *   D[128 frames x 128 slot columns] += A[128 frames x 32 pixels] * B[32 pixels x 128 slot columns]
 * per strip row, 32-pixel chunk and 128-frame tile, float32-accurate by a TF32 main product plus one bf16 MMA for
 * both correction terms; accumulators in tensor memory (4 frame tiles x 128 columns).  Needs d2 % 4 == 0 and a
 * 16-byte aligned movie.  Tables come from pmd_make_strips_tc, the coefficient images from pmd_pack_strips_tc:
 *   items:  [n_items][12] int32 = (first column c0 (multiple of 4), width / 8, first row, rows, first image chunk,
 *           32-pixel chunks per row, first event, events, bg partial index, first slot_ptr entry, 0, 0)
 *   events: [n][4] int32 = (row, slot, first output column, n comps | kind << 8), ascending per item: after that row
 *           the slot is read and cleared; kind 0 stores finished local columns to z, kind 1 ADDS a partial sum of
 *           background columns to zbg (zbg must be zero on entry)
 *   bimg:   per (item, row, chunk) 32 KB: [128 columns][32 pixels] TF32 part + bf16 pair part, SWIZZLE_128B images
 * Outputs as pmd_project_stream (z rows of local tasks written once; one background partial per strip in zbg).
 * replaces: pmd_loader.py:316-346, 392-414 (v_projection / v_projection_routine) in full. */
int pmd_project_stream_tc(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* items,
                          int64_t n_items, const int32_t* events, const void* bimg, const float* mean,
                          const float* inv_std, float* z, int64_t ldz, float* zbg, int64_t ldzbg, int64_t bg_stride,
                          void* stream);

/* HOST function (every pointer is a HOST pointer): tables of pmd_project_stream_tc.  Strips of G block columns
 * (<= 128 pixels), 32 slots of 4 accumulator columns, tasks = (block, <= 4 components) or 4 background components;
 * g_fixed > 0 forces G.  Outputs (caller allocated): items [cap_items][12], slot_ptr [cap_items*33], tasks
 * [cap_tasks][8] = (first row, first column relative to c0, rows, width, first output column, n comps, kind, 0),
 * events [cap_events][4], counts[8] = (n_items, n_tasks, n_events, image chunks, n_parts, max width / 8, G, 0);
 * counts[0] == 0: geometry not supported.
 * replaces: nothing in the reference (host bookkeeping of the new projection kernel). */
int pmd_make_strips_tc(const int32_t* row_starts, int64_t nbr, const int32_t* col_starts, int64_t nbc, int64_t bh,
                       int64_t bw, int64_t d1, int64_t d2, const int64_t* ranks, const int64_t* col0, int64_t n_bg,
                       int64_t g_fixed, int32_t* items_out, int64_t cap_items, int32_t* slot_ptr_out,
                       int32_t* tasks_out, int64_t cap_tasks, int32_t* events_out, int64_t cap_events, int64_t* counts);

/* Builds the coefficient images of pmd_project_stream_tc on the device: one CTA per (item, row) pair listed in
 * item_of_row [n_rows_total][2] = (item, row - first row of the item).  uvals: block-component values
 * [column][bh*bw] float32, bg: [K][d] float32 dense background rows (may be NULL when there are none).
 * replaces: nothing in the reference (operand packing of the new projection kernel). */
int pmd_pack_strips_tc(const int32_t* items, const int32_t* item_of_row, int64_t n_rows_total, const int32_t* slot_ptr,
                       const int32_t* tasks, const float* uvals, int64_t bpix, const float* bg, int64_t d, int64_t d2,
                       void* bimg, void* stream);

/* K7, movie operand in TENSOR MEMORY (tcgen05 TS form, TMA-fed; sm_100a): the projection of pmd_project_stream_tc with
 *   - column strips that partition every image row exactly (no halo: every movie element is read from HBM once),
 *   - raw movie tiles [128 frames x 32 pixels] fetched by 2-D TMA boxes (cp.async.bulk.tensor) in strip-fastest grid
 *     order, centred and split by converter warps and written straight into tensor memory (A operand of the MMAs),
 *   - N = 96 / 128 / 192 slot columns with 384 / N frame tiles of 128 frames per CTA (chosen by pmd_make_strips_ts),
 *   - 1 / std folded into the coefficient images (pmd_pack_strips_ts); mean may be NULL.
 * Tables come from pmd_make_strips_ts:
 *   items:  [n_items][12] int32 = (first column c0, 32-pixel chunks per row, first row, rows, first image chunk,
 *           first event, events, bg partial index, first slot_ptr entry, 0, n_main = number of full-height items (they come first), 0)
 *   events: [n][4] int32 = (row, slot, first output column, n comps | kind << 8), ascending per item: after that row the
 *           slot is read and cleared; kind 0 stores finished local columns to z, kind 1 ADDS a partial sum of background
 *           columns to zbg (zero on entry), kind 2 atomically ADDS the partial sum of a block shared by two strips to z
 *           (those rows of z must be zero on entry; two commuting contributions: the result is order independent)
 *   bimg:   per (item, row, chunk) 2 N 128 bytes: [N columns][32 pixels] TF32 part + bf16 pair part, SWIZZLE_128B
 * Needs d2 % 4 == 0, a 16-byte aligned movie whose frame pitch AND image-row pitch in bytes are multiples of 16.
 * replaces: pmd_loader.py:316-346, 392-414 (v_projection / v_projection_routine) in full. */
int pmd_project_stream_ts(const void* movie, int dtype, int64_t t, int64_t d2, int64_t d, const int32_t* items,
                          int64_t n_items, const int32_t* events, const void* bimg, int64_t n, const float* mean, float* z,
                          int64_t ldz, float* zbg, int64_t ldzbg, int64_t bg_stride, void* stream);

/* HOST function (every pointer is a HOST pointer): tables of pmd_project_stream_ts.  w_fixed / n_fixed > 0 force the
 * strip width (32, 64, 96, 128 pixels) / the slot columns N, otherwise both are chosen by a cost model.  Outputs (caller
 * allocated): items [cap_items][12], slot_ptr [cap_items*49], tasks [cap_tasks][8] = (first row, first column relative
 * to c0 (may be negative), rows, width, first output column, n comps, kind 0 whole block | 1 background | 2 block shared
 * by two strips, 0), events [cap_events][4], counts[8] = (n_items, n_tasks, n_events, image chunks, n_parts, W, N,
 * frame tiles per CTA); counts[0] == 0: geometry not supported (blocks wider than 129 pixels).
 * replaces: nothing in the reference (host bookkeeping of the new projection kernel). */
int pmd_make_strips_ts(const int32_t* row_starts, int64_t nbr, const int32_t* col_starts, int64_t nbc, int64_t bh,
                       int64_t bw, int64_t d1, int64_t d2, const int64_t* ranks, const int64_t* col0, int64_t n_bg,
                       int64_t w_fixed, int64_t n_fixed, int32_t* items_out, int64_t cap_items, int32_t* slot_ptr_out,
                       int32_t* tasks_out, int64_t cap_tasks, int32_t* events_out, int64_t cap_events, int64_t* counts);

/* Coefficient images of pmd_project_stream_ts on the device: one CTA per (item, row) pair listed in item_of_row
 * [n_rows_total][2] = (item, row - first row of the item).  uvals: block-component values [column][bh*bw] float32,
 * bg: [K][d] float32 dense background rows (may be NULL when there are none), inv_std: [d] or NULL (folded into the
 * coefficients), n: slot columns N of the tables.
 * replaces: nothing in the reference (operand packing of the new projection kernel). */
int pmd_pack_strips_ts(const int32_t* items, const int32_t* item_of_row, int64_t n_rows_total, const int32_t* slot_ptr,
                       const int32_t* tasks, const float* uvals, int64_t bpix, const float* bg, const float* inv_std,
                       int64_t d, int64_t d2, int64_t n, void* bimg, void* stream);

/* K7b  full-movie projection onto dense (background) columns:
 *   z[c][f] += sum_p basis[c][p] * (movie[f][p] - mean[p]) * inv_std[p]     (c < k <= 16)
 * replaces: the same v_projection for the dense background columns appended at
 *           decomposition.py:912-933.  z rows must be zeroed by the caller (atomic accumulation). */
int pmd_project_dense(const void* movie, int dtype, int64_t t, int64_t d, const float* basis, int64_t k,
                      const float* mean, const float* inv_std, float* z, int64_t ldz, void* stream);

/* K9  frame reconstruction  out[n][i] = (sum_j U[pix[i]][j] * c[j][n]) * scale[pix[i]] + shift[pix[i]]
 * replaces: pmdarray.py:132-171 (PMDArray.__getitem__: CSR row crop x dense temporal slice,
 *           un-normalise, frames first).  U as CSR over physical pixel rows (float32 values, int32
 *           column ids); c: [R][n] float32; pix: [npix] int32 physical pixel ids; scale/shift may be
 *           NULL.  out: [n][npix] float32. */
int pmd_reconstruct(const int64_t* indptr, const int32_t* indices, const float* values, const float* c,
                    int64_t n, const int32_t* pix, int64_t npix, const float* scale, const float* shift,
                    float* out, void* stream);

/* float64 variants for the whitening step  G = M^T (U^T U) M  (decomposition.py:974-996): the Gram
 * matrix of the denoised init movie spans > 8 decades, so U M and U^T (U M) are formed in float64.
 * pmd_reconstruct_f64: out[n][i] = sum_j U[pix[i]][j] c[j][n]        (CSR float64 values, c [R][n] float64)
 * pmd_project_cols_f64: z[col][f] = sum_p U[p][col] w[f][p]          (w [m][d] float64; local columns from
 *   uvals64 [n_local][bh*bw] via blk_of_col/starts, then n_cols-n_local dense rows bg64 [.][d]); z [n_cols][m]. */
int pmd_reconstruct_f64(const int64_t* indptr, const int32_t* indices, const double* values, const double* c,
                        int64_t n, const int32_t* pix, int64_t npix, double* out, void* stream);
int pmd_project_cols_f64(const double* w, int64_t m, int64_t d2, int64_t d, const int32_t* starts, int64_t bh,
                         int64_t bw, const int32_t* blk_of_col, const int64_t* col0, int64_t n_local,
                         const double* uvals64, const double* bg64, int64_t n_cols, double* z, void* stream);

/* block-sparse Gram of the local columns of U (float64), written directly as canonical CSR:
 * for every ordered pair p = (b1, b2) of blocks whose windows overlap, the dense tile
 *   G[c1][c2] = sum over the overlap of  uvals64[col0[b1]+c1][.] * uvals64[col0[b2]+c2][.]
 * goes to vals[rowptr[col0[b1]+c1] + pair_rowoff[p] + c2] with column id col0[b2]+c2 in cols[...].
 * replaces: the scipy.sparse product u.T.dot(u) implied by decomposition.py:974-981.
 * pairs [n_pairs][2] int32 sorted by (b1, b2); pair_rowoff[p] = number of entries that precede tile p in each of
 * its rows; rowptr [n_local+1] (all built by the host from ranks, localmd_b200/decomposition.py: SparseU.gram). */
int pmd_utu_pairs(const int32_t* pairs, int64_t n_pairs, const int64_t* pair_rowoff, const int32_t* starts,
                  int64_t bh, int64_t bw, const int32_t* ranks, const int64_t* col0, const double* uvals64,
                  const int64_t* rowptr, double* vals, int32_t* cols, void* stream);

/* Z_loc = (U_loc^T U_loc) X in float64 from the dense tiles written by pmd_utu_pairs, on the FP64 tensor cores: one CTA per
 * (block b1, 256 columns of X) walks the pairs seg_ptr[b1] .. seg_ptr[b1 + 1] - 1 of the sorted pair list (seg_ptr [nb + 1]
 * int32, host built from the block grid).  x [n_local][ldx], z [n_local][ldz] (rows of blocks with rank 0 do not exist).
 * replaces: the products with u.T.dot(u) at decomposition.py:974-981 (local x local part). */
int pmd_utu_apply_tiles(const int32_t* pairs, const int32_t* seg_ptr, int64_t nb, const int64_t* pair_rowoff,
                        const int32_t* ranks, const int64_t* col0, const int64_t* rowptr, const double* vals,
                        const double* x, int64_t ldx, int64_t m, double* z, int64_t ldz, void* stream);

/* host-side tables of pmd_utu_pairs (no device work): pair_rowoff [n_pairs] and rowptr [sum(ranks) + 1] from the block
 * pairs (sorted by b1) and the kept ranks [nb].  Plain C++ so that the Python driver can run it on a worker thread
 * without holding the interpreter lock. */
int pmd_utu_host_tables(const int32_t* pairs, int64_t n_pairs, const int64_t* ranks, int64_t nb, int64_t* pair_rowoff,
                        int64_t* rowptr);

/* Operand preparation of the float32-accurate tensor-core GEMMs of the mixing step (pmd_loader.py:411-412,
 * dense @ (sparse @ Yc); decomposition.py:1090-1099 back-multiplications):  a b ~= a_hi b_hi (TF32 GEMM) +
 * [a_lo | a_hi][b_hi ; b_lo] (one bf16 GEMM of twice the depth).  In one pass x [rows][ldx] is overwritten by its TF32-exact
 * part hi and the bf16 image `pair` is written:
 *   mode 0 (right operand, depth = rows): pair [2 rows][ldp]:  pair[r][c] = bf16(hi), pair[rows + r][c] = bf16(lo)
 *   mode 1 (left operand,  depth = cols): pair [rows][ldp >= 2 cols]:  pair[r][c] = bf16(lo), pair[r][cols + c] = bf16(hi)
 * cols, ldx, ldp multiples of 4. */
int pmd_split_tf32_bf16(float* x, int64_t rows, int64_t cols, int64_t ldx, void* pair, int64_t ldp, int mode, void* stream);

/* Large SYMMETRIC float64 product on the FP64 tensor cores (mma.sync m8n8k4):  C[i][j] = sum_k A(i, k) B(j, k), n x n,
 * for a product the caller knows to be symmetric (b == NULL: B = A, the Gram A A^T).  Only the 128 x 128 tiles on and
 * above the diagonal are computed; both triangles of C are written from them, so C is exactly symmetric.
 *   layout 0: A(i, k) = a[i * lda + k], B(j, k) = b[j * ldb + k]   (rows with the inner dimension contiguous)
 *   layout 1: A(i, k) = a[k * lda + i], B(j, k) = b[k * ldb + j]   (C = A^T B for row-major [k_len][n] operands)
 * a_dtype / b_dtype: PMD_F32 or PMD_F64 (float32 operands are converted while they are staged: exact).  Supported:
 * layout 0 f32/f32 and f64/f64, layout 1 f32/f64 and f64/f64.  The inner dimension is cut into `splits` chunks (1..64)
 * whose partial tiles go to `work` (splits * nt (nt + 1) / 2 * 128 * 128 doubles, nt = ceil(n / 128)) and are added in
 * ascending order (deterministic).
 * replaces: the Gram of fewer_rows_svd_routine (decomposition.py:1063-1071) and M^T (U^T U M) of
 *           compute_lowrank_factorized_svd (decomposition.py:974-983). */
int pmd_sym_product_f64(const void* a, int a_dtype, int64_t lda, const void* b, int b_dtype, int64_t ldb, int layout,
                        int64_t n, int64_t k_len, int64_t splits, double* work, double* c, void* stream);

/* CSR export of U straight from the block-component form: canonical CSR (rows = pixels, ascending columns, exact zeros
 * dropped) in two passes of one thread per pixel.  Block grid = row_starts [n_br] x col_starts [n_bc] (ascending; block
 * b = ri * n_bc + ci), kept ranks [nb], first column col0 [nb], weighted values uvals [(col0[b] + c) * bh*bw + q] float64,
 * dense background rows bg [K][d] float32 as columns n_local .. n_local + K - 1.  Two row numberings are produced at once:
 * "rel" (row id = row_ids[p], e.g. the reference's order="F" pixel numbering; row_ids NULL = physical) with float64 values
 * and "phys" (row id = p = i * d2 + j) with float32 values (what pmd_reconstruct reads).
 *   fill = 0: counts_rel [d], counts_phys [d] = entries per row (the caller turns them into indptr by a prefix sum)
 *   fill = 1: cols / vals written at indptr_rel[row] .. / indptr_phys[row] ..
 * replaces: decomposition.py:811-857 (coo triplets -> csr) and 912-933 (aggregate_decomposition: background columns). */
int pmd_export_csr(const double* uvals, const float* bg, int64_t K, int64_t d1, int64_t d2, const int32_t* row_starts,
                   int64_t n_br, const int32_t* col_starts, int64_t n_bc, int64_t bh, int64_t bw, const int32_t* ranks,
                   const int64_t* col0, int64_t n_local, const int64_t* row_ids, int fill, int64_t* counts_rel,
                   int64_t* counts_phys, const int64_t* indptr_rel, const int64_t* indptr_phys, int32_t* cols_rel,
                   double* vals_rel, int32_t* cols_phys, float* vals_phys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMD_SM100_H */
