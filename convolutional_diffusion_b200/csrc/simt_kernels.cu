// Exact fp32 SIMT kernels: unified masked-softmax evaluation of LS / ELS / bbELS, bank preparation
// (strip8 packing, patch norms), the log-sum-exp merge, the score epilogue and the DDIM update.
//
// Reference behaviour restated here (never copied): /root/reference/src/utils/idealscore.py
//   LS    :497-557   same-location candidates, k x k window of squared differences zero-filled at borders
//   ELS   :397-473   circular-padded query patches vs every valid (un-padded) patch of every image
//   bbELS :156-372   zero-padded query; per axis a border query only sees candidates at its own
//                    coordinate, an interior query sees every interior coordinate
// All three are  mu = softmax_{cands}( -||q - a p||^2 / (2 beta) + logw_n ) . centre(p).
#include "common.cuh"
#include "../../include/cdscore.h"

namespace {

struct SimtParams {
  int kind, pad, B, C, H, W, k, splits, region;
  long long n_sel;
  const float* x;
  const float* beta;
  const float* images;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc;
};

constexpr int SIMT_THREADS = 128;

template <int C>
__global__ void __launch_bounds__(SIMT_THREADS) partials_simt_kernel(SimtParams p) {
  extern __shared__ float smem[];
  const int H = p.H, W = p.W, k = p.k, d = k / 2;
  const int Hp = H + 2 * d, Wp = W + 2 * d, plane = Hp * Wp;
  float* xs = smem;               // [C][Hp][Wp] padded query
  float* ts = smem + C * plane;   // [C][Hp][Wp] zero-padded bank image
  const int b = blockIdx.z, split = blockIdx.y, tid = threadIdx.x;
  const int pix = blockIdx.x * SIMT_THREADS + tid;
  bool active = pix < H * W;
  const int i = active ? pix / W : 0, j = active ? pix % W : 0;
  if (p.region != 0) {   // 1: only centre pixels, 2: only border pixels (bbELS centre runs on the tensor cores)
    const bool centre = i >= d && i < H - d && j >= d && j < W - d;
    active = active && ((p.region == 1) == centre);
  }

  const float beta = p.beta[b];
  const float a = sqrtf(1.f - beta);
  const float sc = -CDS_LOG2E / (2.f * beta);
  const float* xb = p.x + (size_t)b * C * H * W;

  for (int e = tid; e < C * plane; e += SIMT_THREADS) {
    int c = e / plane, r = e % plane, y = r / Wp - d, xx = r % Wp - d;
    float v = 0.f;
    if (p.pad == CDS_PAD_CIRCULAR) {
      y = (y % H + H) % H;
      xx = (xx % W + W) % W;
      v = xb[(c * H + y) * W + xx];
    } else if (y >= 0 && y < H && xx >= 0 && xx < W) {
      v = xb[(c * H + y) * W + xx];
    }
    xs[e] = v;
    ts[e] = 0.f;
  }

  // candidate ranges in image coordinates (centre of the candidate patch)
  int u0, u1, v0, v1;
  if (p.kind == CDS_KIND_LS) {
    u0 = i; u1 = i + 1; v0 = j; v1 = j + 1;
  } else if (p.kind == CDS_KIND_ELS) {
    u0 = d; u1 = H - d; v0 = d; v1 = W - d;
  } else {
    const bool rb = (i < d) || (i >= H - d), cb = (j < d) || (j >= W - d);
    u0 = rb ? i : d; u1 = rb ? i + 1 : H - d;
    v0 = cb ? j : d; v1 = cb ? j + 1 : W - d;
  }

  Softmax2<C> sm;
  sm.init();
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  for (long long n = n0; n < n1; ++n) {
    __syncthreads();
    const float* img = p.images + (size_t)p.idx[n] * C * H * W;
    for (int e = tid; e < C * H * W; e += SIMT_THREADS) {
      int c = e / (H * W), r = e % (H * W), y = r / W, xx = r % W;
      ts[c * plane + (y + d) * Wp + xx + d] = __ldg(img + e);
    }
    __syncthreads();
    if (!active) continue;
    const float lw = p.logw[n] * CDS_LOG2E;
    for (int u = u0; u < u1; ++u) {
      for (int v = v0; v < v1; ++v) {
        float dist = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float* xr = xs + c * plane + i * Wp + j;             // top-left of the query patch
          // padded coordinates: image pixel (y,x) sits at (y+d, x+d), so the patch centred at image (u,v)
          // starts at padded (u, v)
          const float* tr = ts + c * plane + u * Wp + v;
          // two-level summation (row sums, then their sum): the rounding error of one running fp32 sum over k*k*C terms
          // reached the 1e-3 budget on mu at k = 35 (3675 terms)
          float cs = 0.f;
          for (int dy = 0; dy < k; ++dy) {
            float rs = 0.f;
            for (int dx = 0; dx < k; ++dx) {
              float df = fmaf(-a, tr[dy * Wp + dx], xr[dy * Wp + dx]);
              rs = fmaf(df, df, rs);
            }
            cs += rs;
          }
          dist += cs;
        }
        float val[C];
#pragma unroll
        for (int c = 0; c < C; ++c) val[c] = ts[c * plane + (u + d) * Wp + v + d];
        sm.push(fmaf(dist, sc, lw), val);
      }
    }
  }
  if (active) {
    const int HW = H * W;
    const size_t o = ((size_t)split * p.B + b) * HW + pix;
    p.m[o] = sm.m;
    p.l[o] = sm.l;
#pragma unroll
    for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b) * C + c) * HW + pix] = sm.acc[c];
  }
}

// Vector-Jacobian product of mu with respect to x in the same unified form (the reference callers differentiate the score
// modules with autograd: src/utils/exterior_derivative.py:68-79, scripts/analyze_exterior_derivative.py:171-183).  With
// normalised weights w_p = exp2(t_p - m)/l of the FINAL merged state and h_p = sum_c g_c (v_p[c] - mu_c),
//   d mu_c / d q_e = (a/beta) sum_p w_p (v_p[c] - mu_c) p_e        (the q_e term cancels because sum_p w_p (v_p - mu) = 0)
// so the gradient that query pixel (i,j) sends to the patch element e = (c', dy, dx) of its padded patch is
//   G[e] = (a/beta) sum_p w_p h_p p_e,
// scattered to x[c', i+dy-d, j+dx-d] (wrapped for circular padding, dropped where the patch element is zero padding).
// One thread per query pixel; per bank image and candidate row u: (1) coefficients w_p h_p of the row's candidates from the
// exact distances, kept in registers; (2) for every patch row (c', dy) the correlation of those coefficients with the bank
// row, accumulated in a shared-memory array G[e][thread].  The caller adds -g/beta and the factor a/beta of the score.
struct VjpParams {
  int kind, pad, B, C, H, W, k, splits, tpb;
  long long n_sel;
  const float* x;
  const float* beta;
  const float* images;
  const int32_t* idx;
  const float* logw;
  const float *m, *l, *mu, *g;
  float* grad;
};

constexpr int VJP_MAXROW = 64;     // candidates per row kept in registers (images up to 64 + ... pixels wide)

template <int C>
__global__ void __launch_bounds__(128) score_vjp_simt_kernel(VjpParams p) {
  extern __shared__ float smem[];
  const int H = p.H, W = p.W, k = p.k, d = k / 2, TPB = p.tpb;
  const int Hp = H + 2 * d, Wp = W + 2 * d, plane = Hp * Wp, Dk = C * k * k;
  float* xs = smem;               // [C][Hp][Wp] padded query
  float* ts = smem + C * plane;   // [C][Hp][Wp] zero-padded bank image
  float* G = ts + C * plane;      // [Dk][TPB]
  const int b = blockIdx.z, split = blockIdx.y, tid = threadIdx.x;
  const int pix = blockIdx.x * TPB + tid;
  const bool active = pix < H * W;
  const int i = active ? pix / W : 0, j = active ? pix % W : 0;
  const float beta = p.beta[b];
  const float a = sqrtf(1.f - beta);
  const float sc = -CDS_LOG2E / (2.f * beta);
  const float* xb = p.x + (size_t)b * C * H * W;
  for (int e = tid; e < C * plane; e += TPB) {
    int c = e / plane, r = e % plane, y = r / Wp - d, xx = r % Wp - d;
    float v = 0.f;
    if (p.pad == CDS_PAD_CIRCULAR) {
      y = (y % H + H) % H;
      xx = (xx % W + W) % W;
      v = xb[(c * H + y) * W + xx];
    } else if (y >= 0 && y < H && xx >= 0 && xx < W) {
      v = xb[(c * H + y) * W + xx];
    }
    xs[e] = v;
    ts[e] = 0.f;
  }
  for (int e = tid; e < Dk * TPB; e += TPB) G[e] = 0.f;
  int u0, u1, v0, v1;
  if (p.kind == CDS_KIND_LS) {
    u0 = i; u1 = i + 1; v0 = j; v1 = j + 1;
  } else if (p.kind == CDS_KIND_ELS) {
    u0 = d; u1 = H - d; v0 = d; v1 = W - d;
  } else {
    const bool rb = (i < d) || (i >= H - d), cb = (j < d) || (j >= W - d);
    u0 = rb ? i : d; u1 = rb ? i + 1 : H - d;
    v0 = cb ? j : d; v1 = cb ? j + 1 : W - d;
  }
  const int HW = H * W;
  float mq = 0.f, inv_l = 0.f, muq[C], gq[C];
  if (active) {
    mq = p.m[(size_t)b * HW + pix];
    inv_l = 1.f / p.l[(size_t)b * HW + pix];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      muq[c] = p.mu[((size_t)b * C + c) * HW + pix];
      gq[c] = p.g[((size_t)b * C + c) * HW + pix];
    }
  }
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  for (long long n = n0; n < n1; ++n) {
    __syncthreads();
    const float* img = p.images + (size_t)p.idx[n] * C * H * W;
    for (int e = tid; e < C * H * W; e += TPB) {
      int c = e / (H * W), r = e % (H * W), y = r / W, xx = r % W;
      ts[c * plane + (y + d) * Wp + xx + d] = __ldg(img + e);
    }
    __syncthreads();
    if (!active) continue;
    const float lw = p.logw[n] * CDS_LOG2E;
    for (int u = u0; u < u1; ++u) {
      float coef[VJP_MAXROW];
      // (1) w_p h_p of the candidates (u, v0..v1)
      for (int v = v0; v < v1; ++v) {
        float dist = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float* xr = xs + c * plane + i * Wp + j;
          const float* tr = ts + c * plane + u * Wp + v;
          float cs = 0.f;
          for (int dy = 0; dy < k; ++dy) {
            float rs = 0.f;
            for (int dx = 0; dx < k; ++dx) {
              float df = fmaf(-a, tr[dy * Wp + dx], xr[dy * Wp + dx]);
              rs = fmaf(df, df, rs);
            }
            cs += rs;
          }
          dist += cs;
        }
        float h = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) h = fmaf(gq[c], ts[c * plane + (u + d) * Wp + v + d] - muq[c], h);
        coef[v - v0] = exp2f(fmaf(dist, sc, lw) - mq) * inv_l * h;
      }
      // (2) G[c', dy, dx] += sum_v coef[v] * T[c'][u - d + dy][v - d + dx]   (padded coordinates: ts[..][u + dy][v + dx])
      for (int c = 0; c < C; ++c)
        for (int dy = 0; dy < k; ++dy) {
          const float* tr = ts + c * plane + (u + dy) * Wp;
          for (int dx = 0; dx < k; ++dx) {
            float s2 = 0.f;
            for (int v = v0; v < v1; ++v) s2 = fmaf(coef[v - v0], tr[v + dx], s2);
            G[((c * k + dy) * k + dx) * TPB + tid] += s2;
          }
        }
    }
  }
  if (!active) return;
  const float ab = a / beta;
  float* gb = p.grad + (size_t)b * C * HW;
  for (int c = 0; c < C; ++c)
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        int y = i + dy - d, xx = j + dx - d;
        if (p.pad == CDS_PAD_CIRCULAR) {
          y = (y % H + H) % H;
          xx = (xx % W + W) % W;
        } else if (y < 0 || y >= H || xx < 0 || xx >= W) {
          continue;                    // zero padding: a constant, no gradient
        }
        atomicAdd(gb + (c * H + y) * W + xx, ab * G[((c * k + dy) * k + dx) * TPB + tid]);
      }
}

__global__ void pack_strip8_kernel(const float* __restrict__ images, long long N, int C, int H, int W, float scale,
                                   int plane, uint4* __restrict__ out) {
  const long long total = N * C * H * W;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int xx = g % W;
    const int u = (g / W) % H;
    const long long nc = g / ((long long)W * H);
    const float* src = images + nc * H * W;
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v;
      if (plane == 2) v = (xx + e < W) ? src[u * W + xx + e] * scale : 0.f;     // rows8: horizontal strip
      else v = (u + e < H) ? src[(u + e) * W + xx] * scale : 0.f;
      __half hi = __float2half_rn(v);
      h[e] = plane == 1 ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    out[g] = *reinterpret_cast<uint4*>(h);
  }
}

__global__ void patch_norms_kernel(const float* __restrict__ images, long long N, int C, int H, int W, int k,
                                   float* __restrict__ out) {
  const int Ph = H - k + 1, Pw = W - k + 1;
  const long long total = N * Ph * Pw;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int v = g % Pw, u = (g / Pw) % Ph;
    const long long n = g / ((long long)Pw * Ph);
    const float* img = images + n * C * H * W;
    float s = 0.f;
    for (int c = 0; c < C; ++c)
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
          float t = img[(c * H + u + dy) * W + v + dx];
          s = fmaf(t, t, s);
        }
    out[g] = s;
  }
}

// log-sum-exp merge of S slices: 32 pixels x 8 slice groups per block; each group folds its slices online, the
// groups are then merged through shared memory (S reaches ~300 for the LS kernel, so a serial loop would be latency bound)
constexpr int CMB_G = 8;
__global__ void combine_kernel(const float* __restrict__ m, const float* __restrict__ l, const float* __restrict__ acc,
                               int S, int B, int C, int HW, size_t stride_ml, size_t stride_acc, float* m_out,
                               float* l_out, float* acc_out) {
  __shared__ float sh[CMB_G][32][10];
  const int lane = threadIdx.x, grp = threadIdx.y;
  const int g = blockIdx.x * 32 + lane;
  const bool on = g < B * HW;
  const int b = on ? g / HW : 0, pix = on ? g % HW : 0;
  float M = -INFINITY, L = 0.f, A[8];
  for (int c = 0; c < C; ++c) A[c] = 0.f;
  if (on) {
    for (int s = grp; s < S; s += CMB_G) {
      const size_t o = (size_t)s * stride_ml + (size_t)b * HW + pix;
      const float ms = m[o];
      if (ms == -INFINITY) continue;
      const float Mn = fmaxf(M, ms);
      const float w0 = exp2f(M - Mn), w1 = exp2f(ms - Mn);
      L = L * w0 + l[o] * w1;
      for (int c = 0; c < C; ++c) A[c] = A[c] * w0 + acc[(size_t)s * stride_acc + ((size_t)b * C + c) * HW + pix] * w1;
      M = Mn;
    }
  }
  sh[grp][lane][0] = M;
  sh[grp][lane][1] = L;
  for (int c = 0; c < C; ++c) sh[grp][lane][2 + c] = A[c];
  __syncthreads();
  if (grp == 0 && on) {
    float Mt = -INFINITY;
    for (int q = 0; q < CMB_G; ++q) Mt = fmaxf(Mt, sh[q][lane][0]);
    float Lt = 0.f, At[8];
    for (int c = 0; c < C; ++c) At[c] = 0.f;
    for (int q = 0; q < CMB_G; ++q) {
      const float mq = sh[q][lane][0];
      const float w = (mq == -INFINITY) ? 0.f : exp2f(mq - Mt);
      Lt = fmaf(sh[q][lane][1], w, Lt);
      for (int c = 0; c < C; ++c) At[c] = fmaf(sh[q][lane][2 + c], w, At[c]);
    }
    m_out[(size_t)b * HW + pix] = Mt;
    l_out[(size_t)b * HW + pix] = Lt;
    for (int c = 0; c < C; ++c) acc_out[((size_t)b * C + c) * HW + pix] = At[c];
  }
}

__global__ void finalize_kernel(const float* __restrict__ x, const float* __restrict__ beta, const float* __restrict__ l,
                                const float* __restrict__ acc, int B, int C, int H, int W, int region, int d,
                                float* __restrict__ mu, float* __restrict__ score) {
  const int HW = H * W;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B * HW) return;
  const int b = g / HW, pix = g % HW, i = pix / W, j = pix % W;
  if (region != 0) {
    const bool rb = i < d || i >= H - d, cb = j < d || j >= W - d;
    const bool centre = !rb && !cb, corner = rb && cb;
    const bool in = region == 1 ? centre : region == 2 ? !centre : region == 3 ? corner : (!centre && !corner);
    if (!in) return;
  }
  const float bt = beta[b], a = sqrtf(1.f - bt);
  const float inv = 1.f / l[(size_t)b * HW + pix];
  for (int c = 0; c < C; ++c) {
    const size_t o = ((size_t)b * C + c) * HW + pix;
    const float mv = acc[o] * inv;
    if (mu) mu[o] = mv;
    if (score) score[o] = -(x[o] - a * mv) / bt;
  }
}

// Philox4x32-10 (Salmon et al., SC'11; the generator behind cuRAND's and PyTorch's CUDA streams), written out so that the
// stochastic sampler needs no library state and can live inside a captured CUDA graph
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// one standard normal per element: counter = (element index, stream offset + step), key = seed; Box-Muller on two words
__device__ __forceinline__ float philox_normal(unsigned long long elem, unsigned long long seed, unsigned long long ctr) {
  uint32_t w[4];
  philox4x32_10((uint32_t)elem, (uint32_t)(elem >> 32), (uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), w);
  const float u1 = ((float)w[0] + 0.5f) * 2.3283064365386963e-10f;      // (0, 1]
  const float u2 = ((float)w[1] + 0.5f) * 2.3283064365386963e-10f;
  return sqrtf(-2.f * __logf(fminf(u1, 1.f))) * __cosf(6.283185307179586f * u2);
}

__global__ void randn_philox_kernel(float* __restrict__ out, long long n, const unsigned long long* __restrict__ so, int step) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g < n) out[g] = philox_normal((unsigned long long)g, so[0], so[1] + (unsigned long long)step);
}

// Fused tail of one evaluation: log-sum-exp merge of the S slices (same scheme as combine_kernel), mu, score, and the
// sampler update in place.  stride_ml / stride_acc select the layout (separate arrays or all-gathered packed records).
__global__ void finish_kernel(const float* __restrict__ m, const float* __restrict__ l, const float* __restrict__ acc, int S,
                              int B, int C, int H, int W, size_t stride_ml, size_t stride_acc, int region, int d,
                              float* __restrict__ x, const float* __restrict__ beta, float* __restrict__ mu,
                              float* __restrict__ score, const float* __restrict__ cx, const float* __restrict__ cmu,
                              const float* __restrict__ sigma, const float* __restrict__ noise,
                              const unsigned long long* __restrict__ seed_offset, int step) {
  __shared__ float sh[CMB_G][32][10];
  const int HW = H * W;
  const int lane = threadIdx.x, grp = threadIdx.y;
  const int g = blockIdx.x * 32 + lane;
  bool on = g < B * HW;
  const int b = on ? g / HW : 0, pix = on ? g % HW : 0;
  if (on && region != 0) {
    const int i = pix / W, j = pix % W;
    const bool rb = i < d || i >= H - d, cb = j < d || j >= W - d;
    const bool centre = !rb && !cb, corner = rb && cb;
    on = region == 1 ? centre : region == 2 ? !centre : region == 3 ? corner : (!centre && !corner);
  }
  float M = -INFINITY, L = 0.f, A[8];
  for (int c = 0; c < C; ++c) A[c] = 0.f;
  if (on) {
    for (int s = grp; s < S; s += CMB_G) {
      const size_t o = (size_t)s * stride_ml + (size_t)b * HW + pix;
      const float ms = m[o];
      if (ms == -INFINITY) continue;
      const float Mn = fmaxf(M, ms);
      const float w0 = exp2f(M - Mn), w1 = exp2f(ms - Mn);
      L = L * w0 + l[o] * w1;
      for (int c = 0; c < C; ++c) A[c] = A[c] * w0 + acc[(size_t)s * stride_acc + ((size_t)b * C + c) * HW + pix] * w1;
      M = Mn;
    }
  }
  sh[grp][lane][0] = M;
  sh[grp][lane][1] = L;
  for (int c = 0; c < C; ++c) sh[grp][lane][2 + c] = A[c];
  __syncthreads();
  if (grp != 0 || !on) return;
  float Mt = -INFINITY;
  for (int q = 0; q < CMB_G; ++q) Mt = fmaxf(Mt, sh[q][lane][0]);
  float Lt = 0.f, At[8];
  for (int c = 0; c < C; ++c) At[c] = 0.f;
  for (int q = 0; q < CMB_G; ++q) {
    const float mq = sh[q][lane][0];
    const float w = (mq == -INFINITY) ? 0.f : exp2f(mq - Mt);
    Lt = fmaf(sh[q][lane][1], w, Lt);
    for (int c = 0; c < C; ++c) At[c] = fmaf(sh[q][lane][2 + c], w, At[c]);
  }
  const float bt = beta[b], a = sqrtf(1.f - bt), inv = 1.f / Lt;
  for (int c = 0; c < C; ++c) {
    const size_t o = ((size_t)b * C + c) * HW + pix;
    const float mv = At[c] * inv, xv = x[o];
    if (mu) mu[o] = mv;
    if (score) score[o] = -(xv - a * mv) / bt;
    if (cx) {
      float xn = cx[b] * xv + cmu[b] * mv;
      if (sigma) {
        const float z = noise ? noise[o] : philox_normal((unsigned long long)o, seed_offset[0], seed_offset[1] + (unsigned long long)step);
        xn = fmaf(sigma[b], z, xn);
      }
      x[o] = xn;
    }
  }
}

__global__ void ddim_step_kernel(float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ cx,
                                 const float* __restrict__ cmu, int B, long long chw) {
  const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (g >= B * chw) return;
  const int b = g / chw;
  x[g] = cx[b] * x[g] + cmu[b] * mu[g];
}

}  // namespace

extern "C" int cds_partials_simt(int kind, int query_pad, const float* x, int B, int C, int H, int W, int k,
                                 const float* beta, const float* images, const int32_t* idx, const float* logw,
                                 int64_t n_sel, int splits, int region, float* m, float* l, float* acc,
                                 void* stream) {
  CDS_CHECK_ARG(region >= 0 && region <= 2, "cds_partials_simt: bad region %d", region);
  CDS_CHECK_ARG(kind >= 0 && kind <= 2, "cds_partials_simt: bad kind %d", kind);
  CDS_CHECK_ARG(C >= 1 && C <= 4, "cds_partials_simt: C=%d unsupported (1..4)", C);
  CDS_CHECK_ARG(k >= 1 && (k & 1), "cds_partials_simt: k=%d must be odd", k);
  CDS_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && n_sel >= 1 && splits >= 1, "cds_partials_simt: empty problem");
  if (kind == CDS_KIND_ELS) CDS_CHECK_ARG(k <= H && k <= W, "cds_partials_simt: ELS needs k <= H,W");
  if (kind == CDS_KIND_BBELS) CDS_CHECK_ARG(k < H && k < W, "cds_partials_simt: bbELS needs k < H,W (k>=H is LS)");
  if (splits > n_sel) splits = (int)n_sel;
  SimtParams p{kind, query_pad, B, C, H, W, k, splits, region, (long long)n_sel, x, beta, images, idx, logw, m, l, acc};
  const int d = k / 2;
  const size_t smem = (size_t)2 * C * (H + 2 * d) * (W + 2 * d) * sizeof(float);
  CDS_CHECK_ARG(smem <= 227 * 1024, "cds_partials_simt: image too large for shared memory (%zu B)", smem);
  dim3 grid((H * W + SIMT_THREADS - 1) / SIMT_THREADS, splits, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(CC)                                                                                           \
  case CC:                                                                                                   \
    cudaFuncSetAttribute(partials_simt_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    partials_simt_kernel<CC><<<grid, SIMT_THREADS, smem, st>>>(p);                                           \
    break;
  switch (C) {
    LAUNCH(1) LAUNCH(2) LAUNCH(3) LAUNCH(4)
  }
#undef LAUNCH
  CDS_CHECK_LAUNCH("partials_simt_kernel");
  return CDS_OK;
}

extern "C" int cds_pack_strip8(const float* images, int64_t N, int C, int H, int W, float scale, int plane,
                               void* out_f16, void* stream) {
  CDS_CHECK_ARG(N >= 1 && C >= 1 && H >= 1 && W >= 1, "cds_pack_strip8: empty bank");
  CDS_CHECK_ARG(plane >= 0 && plane <= 2, "cds_pack_strip8: plane must be 0 (hi), 1 (residual) or 2 (rows8), got %d", plane);
  const long long total = (long long)N * C * H * W;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  pack_strip8_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, C, H, W, scale, plane, (uint4*)out_f16);
  CDS_CHECK_LAUNCH("pack_strip8_kernel");
  return CDS_OK;
}

extern "C" int cds_patch_norms(const float* images, int64_t N, int C, int H, int W, int k, float* out, void* stream) {
  CDS_CHECK_ARG(k >= 1 && k <= H && k <= W, "cds_patch_norms: bad k=%d", k);
  const long long total = (long long)N * (H - k + 1) * (W - k + 1);
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  patch_norms_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, C, H, W, k, out);
  CDS_CHECK_LAUNCH("patch_norms_kernel");
  return CDS_OK;
}

extern "C" int cds_combine(const float* m, const float* l, const float* acc, int S, int B, int C, int HW, float* m_out,
                           float* l_out, float* acc_out, void* stream) {
  CDS_CHECK_ARG(S >= 1 && C <= 8, "cds_combine: bad S=%d C=%d", S, C);
  const int blocks = (B * HW + 31) / 32;
  combine_kernel<<<blocks, dim3(32, CMB_G), 0, (cudaStream_t)stream>>>(m, l, acc, S, B, C, HW, (size_t)B * HW,
                                                                       (size_t)B * C * HW, m_out, l_out, acc_out);
  CDS_CHECK_LAUNCH("combine_kernel");
  return CDS_OK;
}

extern "C" int cds_combine_packed(const float* packed, int S, int B, int C, int HW, float* m_out, float* l_out,
                                  float* acc_out, void* stream) {
  CDS_CHECK_ARG(S >= 1 && C <= 8, "cds_combine_packed: bad S=%d C=%d", S, C);
  const size_t slice = (size_t)B * (2 + C) * HW;     // one rank's [m | l | acc]
  const int blocks = (B * HW + 31) / 32;
  combine_kernel<<<blocks, dim3(32, CMB_G), 0, (cudaStream_t)stream>>>(packed, packed + (size_t)B * HW,
                                                                       packed + (size_t)2 * B * HW, S, B, C, HW, slice,
                                                                       slice, m_out, l_out, acc_out);
  CDS_CHECK_LAUNCH("combine_kernel(packed)");
  return CDS_OK;
}

extern "C" int cds_finalize(const float* x, const float* beta, const float* m, const float* l, const float* acc, int B,
                            int C, int H, int W, int region, int d, float* mu, float* score, void* stream) {
  (void)m;
  const int threads = 128, blocks = (B * H * W + threads - 1) / threads;
  finalize_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(x, beta, l, acc, B, C, H, W, region, d, mu, score);
  CDS_CHECK_LAUNCH("finalize_kernel");
  return CDS_OK;
}

extern "C" int cds_ddim_step(float* x, const float* mu, const float* c_x, const float* c_mu, int B, int64_t chw,
                             void* stream) {
  const int threads = 256;
  const long long total = (long long)B * chw;
  ddim_step_kernel<<<(int)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(x, mu, c_x, c_mu, B, chw);
  CDS_CHECK_LAUNCH("ddim_step_kernel");
  return CDS_OK;
}

extern "C" int cds_finish(const float* m, const float* l, const float* acc, int packed, int S, int B, int C, int H, int W,
                          int region, int d, float* x, const float* beta, float* mu, float* score, const float* c_x,
                          const float* c_mu, const float* sigma, const float* noise, const uint64_t* seed_offset,
                          int step, void* stream) {
  CDS_CHECK_ARG(S >= 1 && C >= 1 && C <= 8 && B >= 1, "cds_finish: bad S=%d C=%d B=%d", S, C, B);
  CDS_CHECK_ARG((c_x == nullptr) == (c_mu == nullptr), "cds_finish: c_x and c_mu come together");
  CDS_CHECK_ARG(sigma == nullptr || (c_x != nullptr && (noise != nullptr || seed_offset != nullptr)),
                "cds_finish: the stochastic update needs c_x/c_mu and either noise or a seed");
  const int HW = H * W;
  const size_t slice = (size_t)B * (2 + C) * HW;
  const float* mm = m;
  const float* ll = l;
  const float* aa = acc;
  size_t s_ml = (size_t)B * HW, s_acc = (size_t)B * C * HW;
  if (packed) {                       // m points at the first rank's record [m | l | acc]
    ll = m + (size_t)B * HW;
    aa = m + (size_t)2 * B * HW;
    s_ml = s_acc = slice;
  }
  const int blocks = (B * HW + 31) / 32;
  finish_kernel<<<blocks, dim3(32, CMB_G), 0, (cudaStream_t)stream>>>(mm, ll, aa, S, B, C, H, W, s_ml, s_acc, region, d, x, beta,
                                                                      mu, score, c_x, c_mu, sigma, noise,
                                                                      (const unsigned long long*)seed_offset, step);
  CDS_CHECK_LAUNCH("finish_kernel");
  return CDS_OK;
}

extern "C" int cds_randn_philox(float* out, int64_t n, const uint64_t* seed_offset, int step, void* stream) {
  CDS_CHECK_ARG(n >= 1 && seed_offset != nullptr, "cds_randn_philox: bad arguments");
  const int threads = 256;
  randn_philox_kernel<<<(int)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      out, (long long)n, (const unsigned long long*)seed_offset, step);
  CDS_CHECK_LAUNCH("randn_philox_kernel");
  return CDS_OK;
}

extern "C" int cds_score_vjp_simt(int kind, int query_pad, const float* x, int B, int C, int H, int W, int k,
                                  const float* beta, const float* images, const int32_t* idx, const float* logw,
                                  int64_t n_sel, int splits, const float* m, const float* l, const float* mu,
                                  const float* g, float* grad_mu, void* stream) {
  CDS_CHECK_ARG(kind >= 0 && kind <= 2, "cds_score_vjp_simt: bad kind %d", kind);
  CDS_CHECK_ARG(C >= 1 && C <= 4, "cds_score_vjp_simt: C=%d unsupported (1..4)", C);
  CDS_CHECK_ARG(k >= 1 && (k & 1), "cds_score_vjp_simt: k=%d must be odd", k);
  CDS_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && n_sel >= 1 && splits >= 1, "cds_score_vjp_simt: empty problem");
  CDS_CHECK_ARG(W - k + 1 <= VJP_MAXROW || kind == CDS_KIND_LS, "cds_score_vjp_simt: more than %d candidates per row", VJP_MAXROW);
  if (kind == CDS_KIND_ELS) CDS_CHECK_ARG(k <= H && k <= W, "cds_score_vjp_simt: ELS needs k <= H,W");
  if (splits > n_sel) splits = (int)n_sel;
  const int d = k / 2;
  const size_t planes = (size_t)2 * C * (H + 2 * d) * (W + 2 * d) * sizeof(float);
  int tpb = 128;
  while (tpb > 32 && planes + (size_t)C * k * k * tpb * sizeof(float) > 200 * 1024) tpb >>= 1;
  const size_t smem = planes + (size_t)C * k * k * tpb * sizeof(float);
  CDS_CHECK_ARG(smem <= 227 * 1024, "cds_score_vjp_simt: geometry too large for shared memory (%zu B)", smem);
  VjpParams p{kind, query_pad, B, C, H, W, k, splits, tpb, (long long)n_sel, x, beta, images, idx, logw, m, l, mu, g, grad_mu};
  dim3 grid((H * W + tpb - 1) / tpb, splits, B);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(CC)                                                                                              \
  case CC:                                                                                                      \
    cudaFuncSetAttribute(score_vjp_simt_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    score_vjp_simt_kernel<CC><<<grid, tpb, smem, st>>>(p);                                                      \
    break;
  switch (C) {
    LAUNCH(1) LAUNCH(2) LAUNCH(3) LAUNCH(4)
  }
#undef LAUNCH
  CDS_CHECK_LAUNCH("score_vjp_simt_kernel");
  return CDS_OK;
}
