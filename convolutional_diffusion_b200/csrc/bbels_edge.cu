// bbELS edge bands: exact fp32 SIMT kernel.
//
// Reference behaviour restated (never copied): /root/reference/src/utils/idealscore.py:256-288.  A query pixel in
// an edge band (patch crosses exactly one image border) at depth r from that border is compared with the
// zero-padded patches of every bank image centred at the SAME depth r and at every interior position along the band;
// value = the candidate's centre pixel; plain softmax weights.  (Centre pixels run on the tensor cores,
// corners are same-location-only = the LS kernel.)
//
// All four bands are brought to one orientation by a flip / transpose applied to x and to the bank image alike
// (distances are invariant): view[t][s], t = across the band (t < d is the zero padding outside the image),
// s = along the band.  One thread per edge pixel; the query patch row lives in registers (kernel size is a template
// parameter).
#include "common.cuh"
#include "../../include/cdscore.h"

namespace {

struct EdgeParams {
  int B, C, H, k, splits;
  long long n_sel;
  const float* x;
  const float* beta;
  const float* images;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc;
};

// image coordinates of view element (t - d, s) for band 0 top, 1 bottom, 2 left, 3 right (square images)
__device__ __forceinline__ int view_to_pixel(int band, int depth, int s, int H) {
  switch (band) {
    case 0: return depth * H + s;
    case 1: return (H - 1 - depth) * H + s;
    case 2: return s * H + depth;
    default: return s * H + (H - 1 - depth);
  }
}

// Distance in dot-product form, |q|^2 - 2a q.p + a^2 |p|^2 with |q|^2 dropped (constant per query): one FMA per patch element
// instead of two, the truncated-patch norms |p|^2 are computed once per candidate (by the thread that owns the same edge
// pixel) and shared through shared memory.  A thread contracts its query patch row against CH candidates at a time from a
// register-cached bank row segment: K + (K+CH-1)/4 shared-memory load instructions (the candidate row is read as 16-byte
// broadcasts) feed K*CH FMAs, so the loop is FMA bound (K = 17: 25 loads per 272 FMAs; the first version of this kernel ran
// at 20 % of the fp32 peak because 37 scalar loads fed 136 FMAs and the shared-memory pipe takes one load per clock).
template <int K>
__global__ void __launch_bounds__(256) bbels_edge_kernel(EdgeParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int D = K / 2;
  constexpr int VR = D + K - 1;            // view rows: D rows of padding + the K-1 image rows a band patch can touch
  constexpr int CH = K <= 21 ? 16 : 8;     // candidates per pass (register budget: CH + K + K+CH-1 values)
  const int H = p.H, C = p.C, I = H - 2 * D;
  const int HP = (H + CH + 7) & ~3;        // padded row length of the bank view: a CH-wide pass may read past the row end
  float* xv = smem;                        // [C][VR][H]
  float* tv = smem + ((C * VR * H + 3) & ~3);       // [C][VR][HP], 16-byte aligned rows
  float* nrm = tv + C * VR * HP;           // [D][I] a^2-free norms of the truncated candidate patches of the staged image
  const int band = blockIdx.x, split = blockIdx.y, b = blockIdx.z, tid = threadIdx.x, nt = blockDim.x;
  const bool active = tid < D * I;
  const int r = active ? tid / I : 0, j = D + (active ? tid % I : 0);
  const float beta = p.beta[b], a = sqrtf(1.f - beta), sc = -CDS_LOG2E / (2.f * beta);
  const float c_dot = -2.f * a * sc, c_nrm = a * a * sc;     // logit = c_dot * q.p + c_nrm * |p|^2 + logw   (log2 units)
  const int HW = H * H;
  const float* xb = p.x + (size_t)b * C * HW;

  for (int e = tid; e < C * VR * H; e += nt) {
    const int c = e / (VR * H), t = (e / H) % VR, s = e % H;
    xv[e] = t >= D ? xb[c * HW + view_to_pixel(band, t - D, s, H)] : 0.f;
  }
  for (int e = tid; e < C * VR * HP; e += nt) tv[e] = 0.f;

  float m = -INFINITY, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  for (long long n = n0; n < n1; ++n) {
    __syncthreads();
    const float* img = p.images + (size_t)p.idx[n] * C * HW;
    for (int e = tid; e < C * (K - 1) * H; e += nt) {
      const int c = e / ((K - 1) * H), t = (e / H) % (K - 1), s = e % H;
      tv[(c * VR + D + t) * HP + s] = __ldg(img + c * HW + view_to_pixel(band, t, s, H));
    }
    __syncthreads();
    if (active) {        // norm of the candidate at (depth r, position j): rows above the image are zero
      float s2 = 0.f;
      for (int c = 0; c < C; ++c)
        for (int dy = D - r; dy < K; ++dy) {
          const float* tr = tv + (c * VR + r + dy) * HP + (j - D);
          float rs = 0.f;
#pragma unroll
          for (int q = 0; q < K; ++q) rs = fmaf(tr[q], tr[q], rs);
          s2 += rs;
        }
      nrm[r * I + (j - D)] = s2;
    }
    __syncthreads();
    if (!active) continue;
    const float lw = p.logw[n] * CDS_LOG2E;
    for (int v0 = D; v0 < H - D; v0 += CH) {
      float dot[CH];
#pragma unroll
      for (int w = 0; w < CH; ++w) dot[w] = 0.f;
      for (int c = 0; c < C; ++c) {
        for (int dy = D - r; dy < K; ++dy) {              // rows above the image are zero in x and T alike
          const float* xr = xv + (c * VR + r + dy) * H + (j - D);
          const float4* tr4 = reinterpret_cast<const float4*>(tv + (c * VR + r + dy) * HP + (v0 - D));   // v0 - D is a multiple of CH
          float xs[K], tt[(K + CH - 1 + 3) & ~3];
#pragma unroll
          for (int q = 0; q < K; ++q) xs[q] = xr[q];
#pragma unroll
          for (int q = 0; q < (K + CH - 1 + 3) / 4; ++q) {
            const float4 v4 = tr4[q];
            tt[4 * q] = v4.x; tt[4 * q + 1] = v4.y; tt[4 * q + 2] = v4.z; tt[4 * q + 3] = v4.w;
          }
          float part[CH];
#pragma unroll
          for (int w = 0; w < CH; ++w) part[w] = 0.f;
#pragma unroll
          for (int q = 0; q < K; ++q)
#pragma unroll
            for (int w = 0; w < CH; ++w) part[w] = fmaf(xs[q], tt[q + w], part[w]);
#pragma unroll
          for (int w = 0; w < CH; ++w) dot[w] += part[w];          // row sums, then their sum: shorter fp32 chains
        }
      }
      // flash-softmax update with the (up to) CH candidates of this pass
      float t[CH], tmax = -INFINITY;
#pragma unroll
      for (int w = 0; w < CH; ++w) {
        const bool on = v0 + w < H - D;
        t[w] = on ? fmaf(dot[w], c_dot, fmaf(nrm[r * I + min(v0 + w - D, I - 1)], c_nrm, lw)) : -INFINITY;
        tmax = fmaxf(tmax, t[w]);
      }
      if (tmax > m) {
        const float s2 = exp2f(m - tmax);
        l *= s2;
        for (int c = 0; c < C; ++c) acc[c] *= s2;
        m = tmax;
      }
#pragma unroll
      for (int w = 0; w < CH; ++w) {
        const float wt = exp2f(t[w] - m);
        l += wt;
        const int vv = min(v0 + w, H - 1);
        for (int c = 0; c < C; ++c) acc[c] = fmaf(wt, tv[(c * VR + D + r) * HP + vv], acc[c]);
      }
    }
  }
  if (active) {
    const int pix = view_to_pixel(band, r, j, H);
    const size_t o = ((size_t)split * p.B + b) * HW + pix;
    p.m[o] = m;
    p.l[o] = l;
    for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b) * C + c) * HW + pix] = acc[c];
  }
}

template <int K>
size_t edge_smem_bytes(int C, int H) {
  constexpr int D = K / 2, VR = D + K - 1, CH = K <= 21 ? 16 : 8;
  const int HP = (H + CH + 7) & ~3, I = H - 2 * D;
  return (size_t)(((C * VR * H + 3) & ~3) + C * VR * HP + D * (I > 0 ? I : 0)) * sizeof(float);
}

template <int K>
int launch_edge(const EdgeParams& p, cudaStream_t st) {
  constexpr int D = K / 2;
  const int I = p.H - 2 * D;
  const int threads = (D * I + 31) / 32 * 32;
  const size_t smem = edge_smem_bytes<K>(p.C, p.H);
  if (threads > 256 || smem > 227 * 1024) return CDS_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(bbels_edge_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return CDS_ERR_CUDA;
  bbels_edge_kernel<K><<<dim3(4, p.splits, p.B), threads, smem, st>>>(p);
  return CDS_OK;
}

}  // namespace

extern "C" int cds_bbels_edge_supported(int C, int H, int W, int k) {
  const int d = k / 2;
  return C >= 1 && C <= 4 && H == W && (k & 1) && k >= 3 && k <= 31 && k < H && d * (H - 2 * d) <= 256 &&
         (size_t)(2 * C * (d + k - 1) * (H + 20) + 256) * 4 <= 227 * 1024;
}

extern "C" int cds_bbels_edge_partials(const float* x, int B, int C, int H, int W, int k, const float* beta,
                                       const float* images, const int32_t* idx, const float* logw, int64_t n_sel,
                                       int splits, float* m, float* l, float* acc, void* stream) {
  if (!cds_bbels_edge_supported(C, H, W, k)) {
    cds_set_error("cds_bbels_edge_partials: unsupported geometry C=%d H=%d W=%d k=%d", C, H, W, k);
    return CDS_ERR_UNSUPPORTED;
  }
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1, "cds_bbels_edge_partials: empty problem");
  if (splits > n_sel) splits = (int)n_sel;
  EdgeParams p{B, C, H, k, splits, (long long)n_sel, x, beta, images, idx, logw, m, l, acc};
  cudaStream_t st = (cudaStream_t)stream;
  int rc = CDS_ERR_UNSUPPORTED;
  switch (k) {
#define KCASE(KK) case KK: rc = launch_edge<KK>(p, st); break;
    KCASE(3) KCASE(5) KCASE(7) KCASE(9) KCASE(11) KCASE(13) KCASE(15) KCASE(17) KCASE(19) KCASE(21) KCASE(23)
    KCASE(25) KCASE(27) KCASE(29) KCASE(31)
#undef KCASE
  }
  if (rc != CDS_OK) {
    cds_set_error("cds_bbels_edge_partials: launch setup failed for k=%d (rc=%d)", k, rc);
    return rc;
  }
  CDS_CHECK_LAUNCH("bbels_edge_kernel");
  return CDS_OK;
}
