/* cdscore.h -- C ABI of the B200-native analytic score machines (LS / ELS / bbELS).
 *
 * This is the drop-in boundary.  The reference (henhen724/convolutional_diffusion) has no FFI: its
 * hot path is the Python callable  module(t, x, label=None, device=None, k=None)  defined in
 * src/utils/idealscore.py (LS :497, ELS :397, bbELS :156) and the DDIM loop :76-118.  Every entry
 * point below replaces one stage of that callable's body; the Python host in
 * convolutional_diffusion_b200/ re-assembles them behind the reference's signatures.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless named *_host
 *   - `stream` is a cudaStream_t passed as void*; every call is stream ordered, never synchronises,
 *     and is CUDA-graph capturable
 *   - return value 0 = ok; otherwise a negative code, text via cds_last_error()
 *   - images are planar fp32 [N][C][H][W] in [-1,1]; x / outputs are planar fp32 [B][C][H][W]
 *   - "partials" are the flash-softmax triple per sample and query pixel:
 *        m   [S][B][H*W]      running max of the logits seen
 *        l   [S][B][H*W]      sum exp(logit - m)
 *        acc [S][B][C][H*W]   sum exp(logit - m) * centre pixel of the candidate patch
 *     S = number of independent bank slices (CTA splits here, GPUs after the all-gather)
 */
#ifndef CDSCORE_H
#define CDSCORE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CDS_KIND_LS    0   /* idealscore.py:476  LocalScoreModule              */
#define CDS_KIND_ELS   1   /* idealscore.py:375  LocalEquivScoreModule         */
#define CDS_KIND_BBELS 2   /* idealscore.py:127  LocalEquivBordersScoreModule  */
#define CDS_PAD_ZEROS    0 /* F.pad(value=0)          idealscore.py:171 */
#define CDS_PAD_CIRCULAR 1 /* F.pad(mode='circular')  idealscore.py:414 */

int         cds_abi_version(void);
const char* cds_last_error(void);
/* SM count / compute capability of the current device (for grid sizing on the host side) */
int         cds_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- bank preparation (once per bank; replaces the per-call DataLoader pass, idealscore.py:184,430,521) */

/* Pack planar fp32 images into the tensor-core streaming layout "strip8":
 *   out[n][c][u][x][8] (fp16), element e = scale * img[n][c][u+e][x]  (0 beyond the last row).
 * One 16-byte granule = an 8-pixel vertical strip, so a k x k patch row block is addressable with
 * 16-byte-stride UMMA descriptors (implicit im2col; patches are never materialised).
 * plane = 0: fp16(scale*v);  plane = 1: fp16 of the rounding residual (second plane for non-8-bit banks);
 * plane = 2: "rows8", the same values as 8-pixel HORIZONTAL strips, element e = scale * img[n][c][u][x+e]
 * (0 beyond the last column), for the mixed K layout of cds_els_partials_umma. */
int cds_pack_strip8(const float* images, int64_t N, int C, int H, int W, float scale, int plane,
                    void* out_f16, void* stream);

/* ||p||^2 of every valid (un-padded) k x k x C patch: out[n][H-k+1][W-k+1]  (idealscore.py:451,243) */
int cds_patch_norms(const float* images, int64_t N, int C, int H, int W, int k, float* out, void* stream);

/* Per-k "norm plane" for the tensor-core kernel: out[n][u][x][8] (fp16), granule = three-way fp16 split of
 * ||p(n,u,x)||^2 (patch with top-left corner (u,x)) laid out (ph,pm,pl,ph,pm,ph,0,0); positions that are not a
 * valid patch carry a large marker so they vanish in the softmax.  It is contracted against the constant
 * -a*scale/2 inside the same UMMA K loop, i.e. the a^2|p|^2 term of idealscore.py:456 costs one K granule. */
int cds_pack_norm_plane(const float* images, int64_t N, int C, int H, int W, int k, void* out_f16, void* stream);

/* ---- score partials (the hot path) */

/* Exact fp32 SIMT evaluation of the unified masked-softmax form for LS / ELS / bbELS.
 * idx[n_sel] selects bank images (class filter / max_samples, resolved on the host),
 * logw[n_sel] is the per-image log-weight (the reference's per-batch mean, idealscore.py:470,553).
 * region: 0 = every query pixel, 1 = only centre pixels (d<=i<H-d, d<=j<W-d), 2 = only border pixels;
 * pixels outside the region are left untouched in m/l/acc.  m is in log2 units. */
int cds_partials_simt(int kind, int query_pad, const float* x, int B, int C, int H, int W, int k,
                      const float* beta, const float* images, const int32_t* idx, const float* logw,
                      int64_t n_sel, int splits, int region, float* m, float* l, float* acc, void* stream);

/* LS (idealscore.py:497-557) as a bank-streaming kernel: the selected images are read from HBM exactly once per
 * launch, squared differences are box-summed separably in shared memory, every thread carries the online softmax of
 * its pixels.  Same partial layout as cds_partials_simt(kind = LS); C in {1,3}, H*W <= 4096. */
int cds_ls_partials(const float* x, int B, int C, int H, int W, int k, const float* beta, const float* images,
                    const int32_t* idx, const float* logw, int64_t n_sel, int splits, float* m, float* l, float* acc,
                    void* stream);

/* The same for images up to 32 x 32 with one warp per image row: horizontal window sums by warp shuffles, one block
 * barrier per round of 8 (C=1) / 4 (C=3) images, and the whole-image window (k >= 2*max(H,W)-1 = the Ideal Score
 * module, idealscore.py:560-636) as a plain sum of rows.  Preferred over cds_ls_partials when supported. */
int cds_ls_rows_supported(int C, int H, int W, int k);
int cds_ls_rows_partials(const float* x, int B, int C, int H, int W, int k, const float* beta, const float* images,
                         const int32_t* idx, const float* logw, int64_t n_sel, int splits, float* m, float* l,
                         float* acc, void* stream);

/* LS on the tensor cores (tcgen05, csrc/ls_umma.cu; single-channel images, 3 <= k < 2*max(H,W)-1): the window sum of x.T_n is
 * the contraction of a banded query matrix A[y][z] = x(z)*[z in win(y)] with the flattened image, M = 128 pixels, K = the
 * image rows their windows touch, N = images (transposed on the fly by cp.async); the window sums of T_n^2 come from a per-k
 * fp32 plane.  Bank side, packed once: cds_pack_flat16 ([n][cds_ls_plane_elems/N] fp16 pixel*scale, same scale as
 * cds_pack_strip8, single exact plane) and cds_pack_ls_norms ([n][cds_ls_norms_elems/N] fp32).  passes as in cds_els_partials_umma.
 * cds_ls_umma_smem_bytes = 0: geometry not supported (use cds_ls_rows_partials / cds_ls_partials). */
int64_t cds_ls_umma_smem_bytes(int C, int H, int W, int k, int passes);
int64_t cds_ls_plane_elems(int64_t N, int H, int W);
int64_t cds_ls_norms_elems(int64_t N, int H, int W);
int cds_pack_flat16(const float* images, int64_t N, int C, int H, int W, float scale, void* out_f16, void* stream);
int cds_pack_ls_norms(const float* images, int64_t N, int C, int H, int W, int k, float* out, void* stream);
int cds_ls_partials_umma(const float* x, int B, int C, int H, int W, int k, const float* beta, const void* flat16,
                         float scale, const float* ls_norms, const int32_t* idx, const float* logw, int64_t n_sel,
                         int splits, int passes, float* m, float* l, float* acc, void* stream);

/* bbELS edge bands (idealscore.py:256-288): queries whose patch crosses exactly one border vs the zero-padded
 * patches at the same depth and every interior position along the band.  Exact fp32; square images, odd k <= 31.
 * Writes the partials of the edge pixels only. */
int cds_bbels_edge_supported(int C, int H, int W, int k);
int cds_bbels_edge_partials(const float* x, int B, int C, int H, int W, int k, const float* beta, const float* images,
                            const int32_t* idx, const float* logw, int64_t n_sel, int splits, float* m, float* l,
                            float* acc, void* stream);

/* The same edge bands on the tensor cores (tcgen05, csrc/bbels_edge_umma.cu): per band ONE contraction with all depths
 * stacked on the query side -- the depth truncation of a patch is a zeroing of query rows, the candidate operand (the k-1
 * image rows under the border as 8-pixel granules across the band) is depth independent.  passes as in
 * cds_els_partials_umma (1 = fp16 query, 2 = fp16 hi + lo); the bank must be one exact fp16 plane (same `scale` as
 * cds_pack_strip8), otherwise use cds_bbels_edge_partials.  Bank side, packed once: cds_pack_edge_plane (k independent,
 * cds_edge_plane_halves fp16 values: [n][band][c][ceil(H/8)][H] granules of 8 pixels across the band) and
 * cds_pack_edge_norms (per k, cds_edge_norms_halves values: squared norms of the truncated patches as K granules).
 * cds_bbels_edge_umma_smem_bytes = 0: geometry not supported.  Writes the partials of the edge pixels only. */
int64_t cds_bbels_edge_umma_smem_bytes(int C, int H, int W, int k, int passes);
int64_t cds_edge_plane_halves(int64_t N, int C, int H);
int64_t cds_edge_norms_halves(int64_t N, int H, int k);
int cds_pack_edge_plane(const float* images, int64_t N, int C, int H, float scale, void* out_f16, void* stream);
int cds_pack_edge_norms(const float* images, int64_t N, int C, int H, int k, void* out_f16, void* stream);
int cds_bbels_edge_partials_umma(const float* x, int B, int C, int H, int W, int k, const float* beta,
                                 const void* edge_plane, float scale, const void* edge_norms, const int32_t* idx,
                                 const float* logw, int64_t n_sel, int splits, int passes, float* m, float* l,
                                 float* acc, void* stream);

/* tcgen05 / TMEM evaluation of ELS (and the bbELS centre region): queries = all H*W pixels of x padded
 * per query_pad, candidates = every valid k x k patch of the selected images, streamed from the strip8
 * bank by bulk-async copies.  passes = 1: fp16 query; 2: fp16 hi+lo query (fp32-grade dot products for
 * 8-bit banks).  bank_lo may be NULL (8-bit-exact bank).  bank_rows may be NULL; when given (the plane = 2 output of
 * cds_pack_strip8: 8-pixel HORIZONTAL strips) and k > 8, k % 8 != 0, the trailing k % 8 patch rows are contracted
 * as horizontal granules instead of one more mostly-empty block of 8 rows (k = 9: 17 UMMAs per tile instead of 28).
 * dbg_dots: optional [B][H*W][P] raw dot dump of the first selected image (tests only, may be NULL).
 * variant selects the epilogue (idealscore.py:456-471: softmax weights and the weighted sum of patch centres):
 *   CDS_ELS_FMA  weights and weighted sums on the FMA pipe (any bank);
 *   CDS_ELS_PV   single sweep, the weights go back to TMEM as an fp16 operand and the weighted sum is a second
 *                contraction O += P.V' on the tensor cores (single-plane banks; error if the geometry is unsupported);
 *   CDS_ELS_AUTO P.V where it is supported and was measured faster (k <= 9), else FMA. */
#define CDS_ELS_AUTO 0
#define CDS_ELS_FMA  1
#define CDS_ELS_PV   2
int cds_els_partials_umma(int query_pad, const float* x, int B, int C, int H, int W, int k,
                          const float* beta, const void* bank_hi, const void* bank_lo, const void* bank_rows,
                          float bank_scale, const void* norm_plane, const int32_t* idx, const float* logw,
                          int64_t n_sel, int splits, int passes, int variant, float* m, float* l, float* acc,
                          float* dbg_dots, void* stream);
/* The same restricted to the query window [qi0, qi0+qrows) x [qj0, qj0+qcols) of x: only the 128-pixel query tiles that
 * cover the window are launched (the bbELS centre region, idealscore.py:218-254, is the window (d,d)+(H-2d, W-2d): 2 tiles
 * instead of 8 for k = 17 on 32x32).  Pixels outside the window are not written, except those that share a tile with it. */
int cds_els_partials_umma_window(int query_pad, const float* x, int B, int C, int H, int W, int k,
                                 const float* beta, const void* bank_hi, const void* bank_lo, const void* bank_rows,
                                 float bank_scale, const void* norm_plane, const int32_t* idx, const float* logw,
                                 int64_t n_sel, int splits, int passes, int variant, int qi0, int qj0, int qrows,
                                 int qcols, float* m, float* l, float* acc, float* dbg_dots, void* stream);
/* 1 when the P.V epilogue supports the geometry */
int cds_els_umma_pv_supported(int C, int H, int W, int k, int passes, int bank_planes);
/* dynamic shared memory the umma kernel needs for this geometry (0 = unsupported geometry) */
int64_t cds_els_umma_smem_bytes(int C, int H, int W, int k, int passes, int bank_planes);

/* Vector-Jacobian product of the denoised estimate: grad_mu[b][c'][y][x'] += sum over query pixels (i,j) and channels c of
 * g[b][c][i][j] * d mu[b][c][i][j] / d x[b][c'][y][x'], exact fp32 SIMT in the unified form of cds_partials_simt (same kind /
 * query_pad / idx / logw).  m, l [B][H*W] and mu [B][C][H*W] are the MERGED results of a cds_partials_simt evaluation of the same
 * inputs (m in log2 units including the |q|^2 term, i.e. not those of the tensor-core kernel).  The reference's callers
 * obtain this product by autograd through the Python modules (src/utils/exterior_derivative.py:68-79); here it is closed form:
 * d mu_c / d q_e = (a/beta) sum_p w_p (v_p[c] - mu_c) p_e.  grad_mu must be zeroed by the caller (atomic adds; slices of the
 * bank -- CTA splits, ranks -- simply accumulate).  The score's gradient is  -g/beta + (a/beta) grad_mu. */
int cds_score_vjp_simt(int kind, int query_pad, const float* x, int B, int C, int H, int W, int k, const float* beta,
                       const float* images, const int32_t* idx, const float* logw, int64_t n_sel, int splits,
                       const float* m, const float* l, const float* mu, const float* g, float* grad_mu, void* stream);

/* ---- merge + epilogue */

/* log-sum-exp merge of S slices into slice 0 of (m_out,l_out,acc_out) (may alias the inputs) */
int cds_combine(const float* m, const float* l, const float* acc, int S, int B, int C, int HW,
                float* m_out, float* l_out, float* acc_out, void* stream);

/* the same merge for S slices that each sit packed as [m | l | acc] (B*(2+C)*HW floats): the layout one rank
 * contributes to the all-gather when the bank is sharded across GPUs, so the exchange is ONE collective */
int cds_combine_packed(const float* packed, int S, int B, int C, int HW, float* m_out, float* l_out, float* acc_out,
                       void* stream);

/* mu = acc/l ; score = -(x - sqrt(1-beta) mu)/beta  (idealscore.py:372,473,557).
 * region: 0 = all pixels, 1 = only pixels with d<=i<H-d and d<=j<W-d (bbELS centre), 2 = the complement,
 * 3 = corners (both coordinates within d of a border), 4 = edge bands (exactly one). */
int cds_finalize(const float* x, const float* beta, const float* m, const float* l, const float* acc,
                 int B, int C, int H, int W, int region, int d, float* mu, float* score, void* stream);

/* Fused tail of one score evaluation: the log-sum-exp merge of cds_combine (packed = 0: S slices in m / l / acc) or
 * cds_combine_packed (packed = 1: m points at S all-gathered records [m | l | acc], l and acc are ignored), mu and score as
 * cds_finalize (same region / d), and -- when c_x is given -- the sampler update in place
 *     x <- c_x[b] x + c_mu[b] mu (+ sigma[b] z),   z ~ N(0,1)
 * i.e. the DDIM step of ScheduledScoreMachine.forward (idealscore.py:101-116) or, with sigma, the DDPM branch of
 * DDIM.sample (src/models.py:48-64) written in terms of the denoised estimate.  z is read from noise [B][C][H*W] when given,
 * else drawn from Philox4x32-10: key = seed_offset[0], counter = (element index, seed_offset[1] + step); seed_offset is a
 * DEVICE buffer of two uint64 so that a captured CUDA graph replays with fresh noise after the host rewrites it.
 * One launch replaces combine -> finalize -> ddim_step (and combine_packed -> finalize -> ddim_step across ranks). */
int cds_finish(const float* m, const float* l, const float* acc, int packed, int S, int B, int C, int H, int W,
               int region, int d, float* x, const float* beta, float* mu, float* score, const float* c_x,
               const float* c_mu, const float* sigma, const float* noise, const uint64_t* seed_offset, int step,
               void* stream);

/* out[i] = the standard normal cds_finish draws for element i at this (seed_offset, step); tests and seeding only */
int cds_randn_philox(float* out, int64_t n, const uint64_t* seed_offset, int step, void* stream);

/* one deterministic DDIM update of ScheduledScoreMachine.forward (idealscore.py:101-116) written in terms
 * of the denoised estimate:  x <- c_x[b]*x + c_mu[b]*mu  */
int cds_ddim_step(float* x, const float* mu, const float* c_x, const float* c_mu, int B, int64_t chw,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif
