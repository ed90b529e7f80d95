// LS score partials: bank-streaming kernel (HBM bound).
//
// Reference behaviour restated (never copied): /root/reference/src/utils/idealscore.py:497-557.  For every pixel
// the candidates are the SAME pixel of every selected bank image; logit_n(i,j) = -(1/2beta) * sum over the k x k
// window around (i,j), zero-filled outside the image, of sum_c (x - a T_n)^2   (+ logw_n).
//
// The bank is read exactly once per launch: a CTA owns a contiguous slice of the selected images and ALL pixels,
// each thread keeps the pixel values of IB images in registers, the squared-difference maps go through shared memory
// for a separable box sum (row pass, column pass), and every thread carries the online-softmax state of its pixels for
// up to SB samples.  Algorithmic bytes per launch = n_sel * C*H*W * 4 (fp32 bank) + O(B*C*H*W).
#include "common.cuh"
#include "../../include/cdscore.h"

namespace {

// images per shared-memory round (amortises the three block barriers); fewer for RGB to bound the registers
__host__ __device__ constexpr int images_per_round(int C) { return C == 1 ? 8 : 4; }
constexpr int SB = 4;      // samples per CTA (softmax states held in registers)
constexpr int MAXPPT = 4;  // pixels per thread (H*W <= 4096)

struct LsParams {
  int B, C, H, W, k, splits, ppt;
  int whole;   // window covers the whole image from every pixel (IS): one block-wide sum per image replaces the box filter
  long long n_sel;
  const float* x;
  const float* beta;
  const float* images;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc;
};

template <int C, int PPT>
__global__ void __launch_bounds__(PPT == 2 ? 512 : 1024) ls_partials_kernel(LsParams p) {
  extern __shared__ float smem[];
  constexpr int IB = images_per_round(C);
  const int H = p.H, W = p.W, k = p.k, d = k / 2, HW = H * W;
  const int Wp = W + 2 * d, Hp = H + 2 * d;
  float* e_map = smem;                       // [IB][H][Wp]   squared differences, zero columns left/right
  float* r_map = smem + IB * H * Wp;         // [IB][Hp][W]   horizontal window sums, zero rows above/below
  const int tid = threadIdx.x, nt = blockDim.x;
  const int split = blockIdx.x, b0 = blockIdx.y * SB;
  const int nb = min(SB, p.B - b0);

  if (!p.whole) {
    for (int e = tid; e < IB * H * Wp; e += nt) e_map[e] = 0.f;
    for (int e = tid; e < IB * Hp * W; e += nt) r_map[e] = 0.f;
  }

  int py[PPT], px[PPT];
  bool act[PPT];
  float xv[SB][PPT][C];
  Softmax2<C> sm[SB][PPT];
  float a_s[SB], sc_s[SB];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int pix = tid + j * nt;
    act[j] = pix < HW;
    py[j] = act[j] ? pix / W : 0;
    px[j] = act[j] ? pix % W : 0;
  }
#pragma unroll
  for (int s = 0; s < SB; ++s) {
    const float beta = s < nb ? p.beta[b0 + s] : 1.f;
    a_s[s] = sqrtf(1.f - beta);
    sc_s[s] = -CDS_LOG2E / (2.f * beta);
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      sm[s][j].init();
#pragma unroll
      for (int c = 0; c < C; ++c)
        xv[s][j][c] = (s < nb && act[j]) ? p.x[((size_t)(b0 + s) * C + c) * HW + py[j] * W + px[j]] : 0.f;
    }
  }
  __syncthreads();

  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  for (long long nbase = n0; nbase < n1; nbase += IB) {
    const int ni = (int)min((long long)IB, n1 - nbase);
    // bank pixels of this round, read once (coalesced) and kept in registers for all samples
    float tv[IB][PPT][C];
    float lw[IB];
#pragma unroll
    for (int ib = 0; ib < IB; ++ib) {
      const bool on = ib < ni;
      const float* img = p.images + (size_t)p.idx[on ? nbase + ib : nbase] * C * HW;
      lw[ib] = on ? p.logw[nbase + ib] * CDS_LOG2E : 0.f;
#pragma unroll
      for (int j = 0; j < PPT; ++j)
#pragma unroll
        for (int c = 0; c < C; ++c) tv[ib][j][c] = (on && act[j]) ? __ldg(img + c * HW + py[j] * W + px[j]) : 0.f;
    }
    // logits of one round for pixel j from its window sums -> online softmax (one max / one rescale per round)
    auto update = [&](int s, int j, const float* box) {
      float t[IB], tmax = -INFINITY;
#pragma unroll
      for (int ib = 0; ib < IB; ++ib) {
        t[ib] = ib < ni ? fmaf(box[ib], sc_s[s], lw[ib]) : -INFINITY;
        tmax = fmaxf(tmax, t[ib]);
      }
      Softmax2<C>& st = sm[s][j];
      if (tmax > st.m) {
        const float scl = exp2f(st.m - tmax);
        st.l *= scl;
#pragma unroll
        for (int c = 0; c < C; ++c) st.acc[c] *= scl;
        st.m = tmax;
      }
#pragma unroll
      for (int ib = 0; ib < IB; ++ib) {
        const float w = exp2f(t[ib] - st.m);
        st.l += w;
#pragma unroll
        for (int c = 0; c < C; ++c) st.acc[c] = fmaf(w, tv[ib][j][c], st.acc[c]);
      }
    };
    if (p.whole) {
      // whole-image window (IS, idealscore.py:560-636): every pixel sees the same distance = block-wide sum
      float* red = smem;                       // [IB][32] per-warp partial sums
      const int warp = tid >> 5, lane = tid & 31, nw = (nt + 31) >> 5;
#pragma unroll
      for (int s = 0; s < SB; ++s) {
        if (s >= nb) break;
        float part[IB];
#pragma unroll
        for (int ib = 0; ib < IB; ++ib) {
          part[ib] = 0.f;
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            if (act[j]) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float df = fmaf(-a_s[s], tv[ib][j][c], xv[s][j][c]);
                part[ib] = fmaf(df, df, part[ib]);
              }
            }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) part[ib] += __shfl_xor_sync(0xffffffffu, part[ib], off);
          if (lane == 0) red[ib * 32 + warp] = part[ib];
        }
        __syncthreads();
        float box[IB];
#pragma unroll
        for (int ib = 0; ib < IB; ++ib) {
          box[ib] = 0.f;
          for (int w = 0; w < nw; ++w) box[ib] += red[ib * 32 + w];      // same order in every thread
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (act[j]) update(s, j, box);
        __syncthreads();
      }
      continue;
    }
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      if (s >= nb) break;
      // (1) squared differences
#pragma unroll
      for (int ib = 0; ib < IB; ++ib)
#pragma unroll
        for (int j = 0; j < PPT; ++j)
          if (act[j]) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float df = fmaf(-a_s[s], tv[ib][j][c], xv[s][j][c]);
              e = fmaf(df, df, e);
            }
            e_map[(ib * H + py[j]) * Wp + px[j] + d] = e;
          }
      __syncthreads();
      // (2) horizontal window sums: the IB images advance together, so IB independent loads are in flight
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (act[j]) {
          const float* row = e_map + py[j] * Wp + px[j];
          float r[IB];
#pragma unroll
          for (int ib = 0; ib < IB; ++ib) r[ib] = 0.f;
#pragma unroll 2
          for (int dx = 0; dx < k; ++dx)
#pragma unroll
            for (int ib = 0; ib < IB; ++ib) r[ib] += row[ib * H * Wp + dx];
#pragma unroll
          for (int ib = 0; ib < IB; ++ib) r_map[(ib * Hp + py[j] + d) * W + px[j]] = r[ib];
        }
      __syncthreads();
      // (3) vertical window sums -> logits -> online softmax
#pragma unroll
      for (int j = 0; j < PPT; ++j)
        if (act[j]) {
          const float* col = r_map + py[j] * W + px[j];
          float box[IB];
#pragma unroll
          for (int ib = 0; ib < IB; ++ib) box[ib] = 0.f;
#pragma unroll 2
          for (int dy = 0; dy < k; ++dy)
#pragma unroll
            for (int ib = 0; ib < IB; ++ib) box[ib] += col[(ib * Hp + dy) * W];
          update(s, j, box);
        }
      __syncthreads();
    }
  }
#pragma unroll
  for (int s = 0; s < SB; ++s) {
    if (s >= nb) break;
#pragma unroll
    for (int j = 0; j < PPT; ++j)
      if (act[j]) {
        const int pix = py[j] * W + px[j];
        const size_t o = ((size_t)split * p.B + b0 + s) * HW + pix;
        p.m[o] = sm[s][j].m;
        p.l[o] = sm[s][j].l;
#pragma unroll
        for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b0 + s) * C + c) * HW + pix] = sm[s][j].acc[c];
      }
  }
}

template <int C>
int launch_ls(const LsParams& p, size_t smem, dim3 grid, int threads, cudaStream_t st) {
  cudaError_t e;
#define GO(P)                                                                                              \
  e = cudaFuncSetAttribute(ls_partials_kernel<C, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e == cudaSuccess) ls_partials_kernel<C, P><<<grid, threads, smem, st>>>(p);
  if (p.ppt == 1) { GO(1) } else if (p.ppt == 2) { GO(2) } else { GO(4) }
#undef GO
  if (e != cudaSuccess) {
    cds_set_error("ls_partials_kernel attribute: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  return CDS_OK;
}

}  // namespace

extern "C" int cds_ls_partials(const float* x, int B, int C, int H, int W, int k, const float* beta,
                               const float* images, const int32_t* idx, const float* logw, int64_t n_sel, int splits,
                               float* m, float* l, float* acc, void* stream) {
  CDS_CHECK_ARG(C == 1 || C == 3, "cds_ls_partials: C=%d unsupported (1 or 3)", C);
  CDS_CHECK_ARG(k >= 1 && (k & 1), "cds_ls_partials: k=%d must be odd", k);
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1, "cds_ls_partials: empty problem");
  CDS_CHECK_ARG(H * W <= 1024 * MAXPPT, "cds_ls_partials: image too large (%dx%d)", H, W);
  if (splits > n_sel) splits = (int)n_sel;
  const int HW = H * W, d = k / 2;
  int ppt = (HW + 1023) / 1024;
  if (ppt == 3) ppt = 4;
  if (C == 3 && ppt < 2 && HW > 512) ppt = 2;          // keep the register footprint of tv[IB][PPT][C] in check
  if (ppt == 2 && (HW + 1) / 2 > 512) ppt = 4;         // the two-pixel instantiation is bounded to 512 threads
  const int threads = ((HW + ppt - 1) / ppt + 31) / 32 * 32;
  const int whole = d >= (H > W ? H : W) - 1 ? 1 : 0;
  LsParams p{B, C, H, W, k, splits, ppt, whole, (long long)n_sel, x, beta, images, idx, logw, m, l, acc};
  const size_t smem = whole ? (size_t)images_per_round(C) * 32 * sizeof(float)
                            : (size_t)images_per_round(C) * (H * (W + 2 * d) + (H + 2 * d) * W) * sizeof(float);
  CDS_CHECK_ARG(smem <= 227 * 1024, "cds_ls_partials: kernel size %d too large for shared memory", k);
  dim3 grid(splits, (B + SB - 1) / SB);
  int rc = C == 1 ? launch_ls<1>(p, smem, grid, threads, (cudaStream_t)stream)
                  : launch_ls<3>(p, smem, grid, threads, (cudaStream_t)stream);
  if (rc != CDS_OK) return rc;
  CDS_CHECK_LAUNCH("ls_partials_kernel");
  return CDS_OK;
}
