// LS score partials, warp-per-row bank-streaming kernel (HBM bound) for images up to 32 x 32.
//
// Reference behaviour restated (never copied): /root/reference/src/utils/idealscore.py:497-557.  For every pixel
// the candidates are the SAME pixel of every selected bank image; logit_n(i,j) = -(1/2beta) * sum over the k x k
// window around (i,j), zero-filled outside the image, of sum_c (x - a T_n)^2   (+ logw_n).
//
// The bank is read exactly once per launch: a CTA owns a contiguous slice of the selected images and ALL pixels.
// One warp per image row (lanes = columns): the squared-difference row is box-summed horizontally with warp shuffles
// (direct for k <= 7, prefix scan otherwise), the row sums of IB images go through a double-buffered shared-memory
// array for the vertical pass (ONE block barrier per round of IB images), and every lane carries the online softmax of
// its pixel for up to SB samples.  A window that covers the whole image from every pixel (k >= 2*max(H,W)-1, i.e. the
// Ideal Score module, idealscore.py:560-636) degenerates to the sum of all rows.
// Algorithmic bytes per launch = n_sel * C*H*W * 4 (fp32 bank) + O(B*C*H*W).
#include "common.cuh"
#include "../../include/cdscore.h"

namespace {

constexpr int SB = 2;      // samples per CTA (softmax states and x live in registers)
#ifndef CDS_LS_IB_C1
#define CDS_LS_IB_C1 8
#endif
#ifndef CDS_LS_IB_C3
#define CDS_LS_IB_C3 4
#endif
#ifndef CDS_LS_MINBLOCKS
#define CDS_LS_MINBLOCKS 1
#endif
__host__ __device__ constexpr int images_per_round(int C) { return C == 1 ? CDS_LS_IB_C1 : CDS_LS_IB_C3; }

struct LsParams {
  int B, C, H, W, k, splits;
  long long n_sel;
  const float* x;
  const float* beta;
  const float* images;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc;
};

// sum of e over lanes [lane-d, lane+d] (values of lanes >= W are zero by construction)
__device__ __forceinline__ float row_box(float e, int d, int lane) {
  if (d <= 3) {
    float r = e;
    for (int q = 1; q <= d; ++q) {
      const float up = __shfl_up_sync(0xffffffffu, e, q), dn = __shfl_down_sync(0xffffffffu, e, q);
      r += (lane >= q ? up : 0.f) + (lane + q < 32 ? dn : 0.f);
    }
    return r;
  }
  float p = e;                                    // inclusive prefix sum
#pragma unroll
  for (int q = 1; q < 32; q <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, p, q);
    if (lane >= q) p += up;
  }
  const int hi = min(lane + d, 31), lo = lane - d - 1;
  const float ph = __shfl_sync(0xffffffffu, p, hi), pl = __shfl_sync(0xffffffffu, p, max(lo, 0));
  return ph - (lo >= 0 ? pl : 0.f);
}

template <int C>
__global__ void __launch_bounds__(1024, CDS_LS_MINBLOCKS) ls_rows_kernel(LsParams p) {
  extern __shared__ float smem[];
  constexpr int IB = images_per_round(C);
  const int H = p.H, W = p.W, k = p.k, HW = H * W;
  const bool whole = k >= 2 * max(H, W) - 1;     // window = whole image from every pixel (IS)
  const int d = whole ? 0 : k / 2;               // vertical halo rows kept zero in shared memory
  const int Hp = H + 2 * d;
  float* hs = smem;                              // [2][IB][Hp][32] horizontal window sums (double buffered)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = warp;                            // one warp per image row
  const int split = blockIdx.x, b0 = blockIdx.y * SB;
  const int nb = min(SB, p.B - b0);
  const bool col_on = lane < W;
  const int dh = whole ? 31 : min(k / 2, 31);    // horizontal half window (clipped to the warp)

  for (int e = threadIdx.x; e < 2 * IB * Hp * 32; e += blockDim.x) hs[e] = 0.f;

  float xv[SB][C], a_s[SB], sc_s[SB];
  Softmax2<C> sm[SB];
#pragma unroll
  for (int s = 0; s < SB; ++s) {
    const float beta = s < nb ? p.beta[b0 + s] : 1.f;
    a_s[s] = sqrtf(1.f - beta);
    sc_s[s] = -CDS_LOG2E / (2.f * beta);
    sm[s].init();
#pragma unroll
    for (int c = 0; c < C; ++c) xv[s][c] = (s < nb && col_on) ? p.x[((size_t)(b0 + s) * C + c) * HW + y * W + lane] : 0.f;
  }
  __syncthreads();

  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  int round = 0;
  for (long long nbase = n0; nbase < n1; nbase += IB) {
    const int ni = (int)min((long long)IB, n1 - nbase);
    // this row of IB bank images: read once (coalesced rows), kept in registers for all samples
    float tv[IB][C], lw[IB];
#pragma unroll
    for (int ib = 0; ib < IB; ++ib) {
      const bool on = ib < ni;
      const long long gi = p.idx[on ? nbase + ib : nbase];
      lw[ib] = on ? p.logw[nbase + ib] * CDS_LOG2E : -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) tv[ib][c] = (on && col_on) ? __ldg(p.images + ((size_t)gi * C + c) * HW + y * W + lane) : 0.f;
    }
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      if (s >= nb) break;
      float* buf = hs + (size_t)((round & 1) * IB) * Hp * 32;
      ++round;
      // (1) squared differences of this row, horizontal window sums by shuffles
#pragma unroll
      for (int ib = 0; ib < IB; ++ib) {
        float e = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float df = fmaf(-a_s[s], tv[ib][c], xv[s][c]);
          e = fmaf(df, df, e);
        }
        e = col_on ? e : 0.f;
        buf[(ib * Hp + y + d) * 32 + lane] = row_box(e, dh, lane);
      }
      __syncthreads();
      // (2) vertical window sums -> logits -> online softmax (one rescale per round).  The other buffer is written in
      // the next round; this one again two rounds from now, after everyone has passed the next barrier.
      float t[IB], tmax = -INFINITY;
#pragma unroll
      for (int ib = 0; ib < IB; ++ib) t[ib] = 0.f;
      if (whole) {
        for (int r = 0; r < H; ++r)
#pragma unroll
          for (int ib = 0; ib < IB; ++ib) t[ib] += buf[(ib * Hp + r) * 32 + lane];
      } else {
        for (int dy = 0; dy < k; ++dy)
#pragma unroll
          for (int ib = 0; ib < IB; ++ib) t[ib] += buf[(ib * Hp + y + dy) * 32 + lane];
      }
#pragma unroll
      for (int ib = 0; ib < IB; ++ib) {
        t[ib] = fmaf(t[ib], sc_s[s], lw[ib]);
        tmax = fmaxf(tmax, t[ib]);
      }
      Softmax2<C>& st = sm[s];
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
        for (int c = 0; c < C; ++c) st.acc[c] = fmaf(w, tv[ib][c], st.acc[c]);
      }
    }
  }
  if (col_on) {
#pragma unroll
    for (int s = 0; s < SB; ++s) {
      if (s >= nb) break;
      const int pix = y * W + lane;
      const size_t o = ((size_t)split * p.B + b0 + s) * HW + pix;
      p.m[o] = sm[s].m;
      p.l[o] = sm[s].l;
#pragma unroll
      for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b0 + s) * C + c) * HW + pix] = sm[s].acc[c];
    }
  }
}

size_t ls_rows_smem_bytes(int C, int H, int W, int k) {
  const bool whole = k >= 2 * (H > W ? H : W) - 1;
  const int d = whole ? 0 : k / 2;
  return (size_t)2 * images_per_round(C) * (H + 2 * d) * 32 * sizeof(float);
}

}  // namespace

extern "C" int cds_ls_rows_supported(int C, int H, int W, int k) {
  return (C == 1 || C == 3) && W <= 32 && H <= 32 && (k & 1) && k >= 1 && ls_rows_smem_bytes(C, H, W, k) <= 200 * 1024;
}

extern "C" int cds_ls_rows_partials(const float* x, int B, int C, int H, int W, int k, const float* beta,
                                    const float* images, const int32_t* idx, const float* logw, int64_t n_sel,
                                    int splits, float* m, float* l, float* acc, void* stream) {
  if (!cds_ls_rows_supported(C, H, W, k)) {
    cds_set_error("cds_ls_rows_partials: unsupported geometry C=%d H=%d W=%d k=%d (C in {1,3}, H,W <= 32, odd k)", C, H, W, k);
    return CDS_ERR_UNSUPPORTED;
  }
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1, "cds_ls_rows_partials: empty problem");
  if (splits > n_sel) splits = (int)n_sel;
  LsParams p{B, C, H, W, k, splits, (long long)n_sel, x, beta, images, idx, logw, m, l, acc};
  const size_t smem = ls_rows_smem_bytes(C, H, W, k);
  dim3 grid(splits, (B + SB - 1) / SB);
  const int threads = 32 * H;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (C == 1) {
    e = cudaFuncSetAttribute(ls_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) ls_rows_kernel<1><<<grid, threads, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(ls_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) ls_rows_kernel<3><<<grid, threads, smem, st>>>(p);
  }
  if (e != cudaSuccess) {
    cds_set_error("ls_rows_kernel attribute: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  CDS_CHECK_LAUNCH("ls_rows_kernel");
  return CDS_OK;
}
