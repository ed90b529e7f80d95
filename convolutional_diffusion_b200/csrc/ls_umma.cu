// LS (same-location k x k window posterior mean) on the tensor cores (tcgen05 UMMA, accumulators in TMEM), C = 1.
//
// Reference behaviour restated (never copied): /root/reference/src/utils/idealscore.py:497-557.  For pixel y the candidates
// are the SAME pixel of every selected bank image n; logit_n(y) = -(1/2beta) sum_{z in win(y)} (x(z) - a T_n(z))^2 + logw_n
// with the window zero-filled outside the image on both sides.  Expanded:
//     sum_win (x - aT)^2 = sum_win x^2  -  2a sum_win x T_n  +  a^2 sum_win T_n^2 .
// The first term is constant per pixel and cancels in the softmax.  The last is a bank-only quantity, packed once per k
// (cds_pack_ls_norms, fp32).  The middle one is the box filter of the product x.T_n -- the expensive part of the SIMT kernel
// (shuffles + a shared-memory round trip per image, pixel and sample) -- and it is linear in T_n:
//     sum_{z in win(y)} x(z) T_n(z) = sum_z A[y][z] T_n(z),   A[y][z] = x(z) * [z in win(y)],
// a banded matrix that depends on the query only.  So S[y][n] = A[y][.] . T_n[.] is ONE contraction with the window selection
// on the query operand (the same move that put the bbELS edge bands on the tensor cores): M = 128 consecutive pixels of the
// flattened image, K = the contiguous slice of the flattened image that their windows touch (rows r0-d .. r1+d), N = images.
//
// The query operand lives in TMEM (lane = pixel, built once per CTA with tcgen05.st: 4 columns per K granule), so shared memory
// is all pipeline.  The candidate operand wants, per K granule (8 consecutive pixels), 8 images interleaved at a 16-byte pitch
// (no-swizzle core matrix); an arbitrary selection idx[] does not come that way, so the band of every image is bulk-copied as it
// lies (one cp.async.bulk per image into a 2-slot raw ring, issued by two fetcher warps) and four warps transpose it shared -> shared into
// [granule][image] rows (16-byte moves, both sides conflict free thanks to a 16-byte pad of either pitch).  [First cut: 16-byte
// cp.async copies straight to the transposed position -- 4 400 cycles to issue one tile, the LSU keeps only ~32 of them in
// flight; the bulk-copy engine has no such limit.]  Warp 8 issues the UMMAs (N = G images, two TMEM accumulator buffers), warps
// 0-3 run the flash-softmax epilogue with TMEM lane = pixel: the value T_n(y) comes from the transposed band, the window norms
// from a region two more fetcher warps fill by bulk copies (one per image, straight into place).  CTA = (M tile, sample) x bank slice;
// the CTAs of one slice run side by side and share the slice's bytes through L2.
#include "umma_common.cuh"

namespace {
using namespace umma;

// warps 0-3 epilogue (even 16-column chunks), 4-7 transposers, 8 MMA issuer, 9-12 band fetchers, 13-16 epilogue (odd chunks;
// warp & 3 = TMEM lane quadrant), 17-20 norm fetchers
constexpr int LS_THREADS = 672;

#ifdef CDS_PROFILE_SWITCHES
__device__ long long g_ls_clk[8][64];
#define LS_STAMP(row, T) do { if ((p.flags & 8) && blockIdx.x == 0 && blockIdx.y == 0 && (T) < 64) g_ls_clk[row][T] = clock64(); } while (0)
#else
#define LS_STAMP(row, T) do { } while (0)
#endif

struct LsUmmaGeom {
  int H, W, k, d, HW, HWp, HWn, MT, passes;
  int KG;            // K granules of the widest band, rounded up to even
  int G;             // images per tile = UMMA N
  int a_col;         // TMEM column of the query operand: passes planes of KG*4 columns behind the two accumulator buffers
  int raw_stride;    // KG*16 + 16: one image's band in the raw ring
  int raw_bytes;     // G * raw_stride
  int b_row;         // G*16 + 16: one K granule of all G images (+ pad)
  int b_bytes, p2_bytes, buf_bytes;
  int smem_total;
};

struct LsUmmaParams {
  LsUmmaGeom g;
  int B, splits;
  long long n_sel;
  const float* x;
  const float* beta;
  const __half* plane;       // [n][HWp] fp16 pixel * scale
  const float* norms;        // [n][HWn] fp32 window sums of T^2, HWn = MT*128
  float scale;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc;
  int flags;                 // CDS_LS_DEBUG (timing experiments only): 1 = no epilogue math, 2 = no UMMAs, 8 = clock stamps
};

// band of M tile mt: first granule-aligned K element and number of granules (even)
__host__ __device__ inline void ls_band(int H, int W, int d, int mt, int& zlo8, int& kg) {
  const int HW = H * W;
  const int y0 = 128 * mt, y1 = min(HW - 1, y0 + 127);
  const int r0 = max(0, y0 / W - d), r1 = min(H - 1, y1 / W + d);
  zlo8 = (r0 * W) & ~7;
  const int zhi = (r1 + 1) * W;
  kg = (zhi - zlo8 + 7) / 8;
  kg = (kg + 1) & ~1;
}

bool make_ls_geom(int C, int H, int W, int k, int passes, LsUmmaGeom& g) {
  if (C != 1 || (k & 1) == 0 || k < 3 || passes < 1 || passes > 2) return false;
  if (H < 4 || W < 4 || H > 64 || W > 64) return false;
  if (k >= 2 * (H > W ? H : W) - 1) return false;          // whole-image window (IS): a plain dot product, SIMT kernel
  g.H = H; g.W = W; g.k = k; g.d = k / 2; g.HW = H * W; g.passes = passes;
  g.HWp = ((g.HW + 7) / 8 + 1) * 8;
  g.MT = (g.HW + 127) / 128;
  g.HWn = g.MT * 128;
  g.KG = 0;
  for (int mt = 0; mt < g.MT; ++mt) {
    int z, kg;
    ls_band(H, W, g.d, mt, z, kg);
    if (z + 8 * kg > g.HWp) return false;
    if (kg > g.KG) g.KG = kg;
  }
  g.raw_stride = g.KG * 16 + 16;
  g.G = 0;
  const int cand[4] = {128, 64, 32, 16};                  // powers of two: 128 / G loader threads per image
  for (int q = 0; q < 4; ++q) {
    const int G = cand[q];
    if (2 * G + passes * g.KG * 4 > TMEM_COLS) continue;
    const int buf = (g.KG * (G * 16 + 16) + G * 512 + 512 + 127) / 128 * 128;
    if (2 * G * g.raw_stride + 2 * buf + 256 <= 227 * 1024) { g.G = G; break; }
  }
  if (g.G == 0) return false;
  g.a_col = 2 * g.G;
  g.raw_bytes = g.G * g.raw_stride;
  g.b_row = g.G * 16 + 16;
  g.b_bytes = g.KG * g.b_row;
  g.p2_bytes = g.G * 512;
  g.buf_bytes = (g.b_bytes + g.p2_bytes + 512 + 127) / 128 * 128;
  g.smem_total = 2 * g.raw_bytes + 2 * g.buf_bytes + 256;
  return true;
}

__global__ void __launch_bounds__(LS_THREADS, 1) ls_umma_kernel(const __grid_constant__ LsUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const LsUmmaGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x % g.MT, b = blockIdx.x / g.MT, split = blockIdx.y;
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  const int n_img = (int)(n1 - n0);
  const int tiles = (n_img + g.G - 1) / g.G;
  const int W = g.W, HW = g.HW, d = g.d, G = g.G;
  int zlo8, KGt;
  ls_band(g.H, W, d, mt, zlo8, KGt);

  uint8_t* sRaw = smem;
  uint8_t* sBuf = smem + 2 * g.raw_bytes;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + 2 * g.raw_bytes + 2 * g.buf_bytes);
  // barriers: rfull[2] (raw band bytes landed), bfull[2] (transposed band + log-weights written), tfull[2] (accumulators complete),
  // done[2] (epilogue finished with buffer / accumulators of that index), pfull[2] (window norms of the tile's pixels landed),
  // rfree[2] (raw slot read by all transposer warps)
  const uint32_t bar_rfull = smem_u32(sBar), bar_bfull = bar_rfull + 16, bar_tfull = bar_rfull + 32, bar_done = bar_rfull + 48;
  const uint32_t bar_pfull = bar_rfull + 64, bar_rfree = bar_rfull + 80;
  uint32_t* sTmemBase = reinterpret_cast<uint32_t*>(sBar + 16);

  const float beta = p.beta[b];
  const float a = sqrtf(1.f - beta);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_rfull + 8 * s, 4);       // the four band fetcher warps (expect_tx each)
      mbar_init(bar_pfull + 8 * s, 4);       // the four norm fetcher warps
      mbar_init(bar_rfree + 8 * s, 4);       // the four transposer warps
      mbar_init(bar_bfull + 8 * s, 4);
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_done + 8 * s, 8);       // the eight epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(sTmemBase), TMEM_COLS);
  // the norms of images past the end of a slice are never fetched; their columns are masked by lw = -inf, which only works on
  // finite values: start from zeros (later tiles leave finite stale data)
  for (int e = tid * 16; e < 2 * g.buf_bytes; e += LS_THREADS * 16) *reinterpret_cast<uint4*>(sBuf + e) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmemBase;

  if (warp < 4) {
    // query operand into TMEM: row yl (= TMEM lane), K step ks = 16 consecutive band elements: A[yl][z] = x(z) if z lies in the
    // k x k window of pixel y = 128 mt + yl, else 0; second plane = fp16 residual of x
    const float* xb = p.x + (size_t)b * HW;
    const int yl = tid, y = 128 * mt + yl;
    const int ry = y / W, cy = y - ry * W;
    const uint32_t arow = tmem_base + ((uint32_t)(warp * 32) << 16) + g.a_col;
    int rz = zlo8 / W, cz = zlo8 - rz * W;
    for (int ks = 0; ks < KGt / 2; ++ks) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int z = zlo8 + 16 * ks + 2 * q + h;
          const bool ok = y < HW && z < HW && abs(rz - ry) <= d && abs(cz - cy) <= d;
          v[h] = ok ? xb[z] : 0.f;
          if (++cz == W) { cz = 0; ++rz; }
        }
        hi[q] = pack_f16x2(v[0], v[1]);
        const __half2 hh = *reinterpret_cast<const __half2*>(&hi[q]);
        lo[q] = pack_f16x2(v[0] - __low2float(hh), v[1] - __high2float(hh));
      }
      tmem_st8(arow + 8 * ks, hi);
      if (g.passes > 1) tmem_st8(arow + g.KG * 4 + 8 * ks, lo);
    }
    tmem_st_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- transposers: raw band [image][granule] -> [granule][image]
    const int tt = tid - 128;
    const int tpi = 128 / G, n = tt / tpi, part = tt - n * tpi;
    auto logw_of = [&](int T) -> float {      // log-weight of this thread's image in tile T, fetched one tile ahead
      return (T < tiles && part == 0 && n < min(G, n_img - T * G)) ? __ldg(p.logw + n0 + (long long)T * G + n) * CDS_LOG2E : -INFINITY;
    };
    float lw_cur = logw_of(0);
    for (int T = 0; T < tiles; ++T) {
      const int s = T & 1;
      const int nv = min(G, n_img - T * G);
      const float lw_next = logw_of(T + 1);
      mbar_wait(bar_rfull + 8 * s, (T >> 1) & 1, 1);
      if (tt == 0) LS_STAMP(0, T);
      mbar_wait(bar_done + 8 * s, ((T >> 1) & 1) ^ 1, 2);         // tile T-2 has left the transposed buffer
      if (tt == 0) LS_STAMP(1, T);
      const uint8_t* raw = sRaw + (size_t)s * g.raw_bytes + (size_t)n * g.raw_stride;
      uint8_t* bc = sBuf + (size_t)s * g.buf_bytes;
      float* lws = reinterpret_cast<float*>(bc + g.b_bytes + g.p2_bytes);
      if (n < nv) {
        int kg = part;
        for (; kg + 5 * tpi < KGt; kg += 6 * tpi) {          // six independent 16-byte moves per round
          uint4 v[6];
#pragma unroll
          for (int u = 0; u < 6; ++u) v[u] = *reinterpret_cast<const uint4*>(raw + (kg + u * tpi) * 16);
#pragma unroll
          for (int u = 0; u < 6; ++u) *reinterpret_cast<uint4*>(bc + (size_t)(kg + u * tpi) * g.b_row + (size_t)n * 16) = v[u];
        }
        for (; kg < KGt; kg += tpi)
          *reinterpret_cast<uint4*>(bc + (size_t)kg * g.b_row + (size_t)n * 16) = *reinterpret_cast<const uint4*>(raw + kg * 16);
      } else {
        for (int kg = part; kg < KGt; kg += tpi) *reinterpret_cast<uint4*>(bc + (size_t)kg * g.b_row + (size_t)n * 16) = make_uint4(0, 0, 0, 0);
      }
      if (part == 0) lws[n] = lw_cur;
      lw_cur = lw_next;
      fence_proxy_async();           // the UMMAs read the band, and the next bulk copy overwrites the raw slot, through the async proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_bfull + 8 * s);
        mbar_arrive(bar_rfree + 8 * s);      // this warp has finished reading raw slot s
      }
      if (tt == 0) LS_STAMP(2, T);
    }
  } else if (warp == 8) {
    // ---------------------------------------------------------------- MMA issuer
    const uint64_t hi128 = desc_hi(128);
    const uint32_t idesc = (1u << 4) | (((uint32_t)G >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b_lbo = ((uint32_t)g.b_row >> 4) << 16;
    const uint32_t a_tmem = tmem_base + g.a_col;
    for (int T = 0; T < tiles; ++T) {
      const int s = T & 1;
      mbar_wait(bar_bfull + 8 * s, (T >> 1) & 1, 3);       // implies done[s] of tile T-2 (the loaders waited for it)
      tc_fence_after();
      if (lane == 0) LS_STAMP(4, T);
      if (elect_one()) {
        const uint32_t b_base = smem_u32(sBuf + (size_t)s * g.buf_bytes) >> 4;
        const uint32_t d_tmem = tmem_base + s * G;
        uint32_t first = 0;
        for (int pa = 0; pa < ((p.flags & 2) ? 0 : g.passes); ++pa)
          for (int ks = 0; ks < KGt / 2; ++ks) {
            const uint32_t blo = (b_base + (uint32_t)((ks * 2 * g.b_row) >> 4)) | b_lbo;
            umma_f16_ts(d_tmem, a_tmem + pa * g.KG * 4 + 8 * ks, hi128 | (uint64_t)blo, idesc, first);
            first = 1u;
          }
        umma_commit(bar_tfull + 8 * s);
      }
      __syncwarp();
      if (lane == 0) LS_STAMP(5, T);
    }
  } else if ((warp >= 9 && warp < 13) || warp >= 17) {
    // ---------------------------------------------------------------- fetchers: one bulk copy per image and tile.  A warp gets a
    // bulk copy out every ~100 cycles, whether its lanes issue one each or one lane loops (both measured), and the window norms of
    // tile T can only be fetched once the epilogue has released the buffer of tile T-2: with two warps per plane (32 copies each)
    // the norms of a tile landed 4 000 cycles after that release and the epilogue waited for them.  Four warps per plane: warps
    // 9-12 the raw bands (slot free once the four transposer warps have read it), warps 17-20 the window norms of the tile's 128
    // pixels, straight into place.
    const bool band = warp < 13;
    const int part = (warp - 9) & 3, per = G / 4;              // images [part*per, part*per + per) of the tile, per <= 32
    const uint32_t bar_free = band ? bar_rfree : bar_done, bar_full = band ? bar_rfull : bar_pfull;
    const uint32_t bytes = band ? (uint32_t)KGt * 16 : 512u;
    const size_t stride = band ? (size_t)g.HWp * 2 : (size_t)g.HWn * 4;
    const uint8_t* src0 = band ? reinterpret_cast<const uint8_t*>(p.plane + zlo8) : reinterpret_cast<const uint8_t*>(p.norms + 128 * mt);
    const uint32_t dst_stride = band ? (uint32_t)g.raw_stride : 512u;
    auto lookup = [&](int T) -> int {
      const int q = part * per + lane;
      return (T < tiles && lane < per && q < min(G, n_img - T * G)) ? p.idx[n0 + (long long)T * G + q] : -1;
    };
    int img0 = lookup(0), img1 = lookup(1);                    // index lookups run two tiles ahead
    for (int T = 0; T < tiles; ++T) {
      const int s = T & 1;
      const int nv = min(G, n_img - T * G);
      const int mine = max(0, min(per, nv - part * per));
      const int img2 = lookup(T + 2);
      mbar_wait(bar_free + 8 * s, ((T >> 1) & 1) ^ 1, 7);
      if (lane == 0) mbar_expect_tx(bar_full + 8 * s, (uint32_t)mine * bytes);
      __syncwarp();
      const uint32_t dst = (band ? smem_u32(sRaw + (size_t)s * g.raw_bytes) : smem_u32(sBuf + (size_t)s * g.buf_bytes + g.b_bytes))
                           + (part * per + lane) * dst_stride;
      if (img0 >= 0) bulk_g2s(dst, src0 + (size_t)img0 * stride, bytes, bar_full + 8 * s);
      if (band && warp == 9 && lane == 0) LS_STAMP(3, T);
      img0 = img1;
      img1 = img2;
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = pixel
    // two warpgroups share every tile: group 0 (warps 0-3) takes the even 16-column chunks, group 1 (warps 13-16) the odd ones; each
    // keeps its own softmax state per pixel, merged once at the end
    const int wg = warp < 4 ? 0 : 1;
    const int yl = (warp & 3) * 32 + lane, y = 128 * mt + yl;
    const float cs = a * CDS_LOG2E / (beta * p.scale), cn = -(1.f - beta) * CDS_LOG2E / (2.f * beta);
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    const int zr = min(max(y - zlo8, 0), KGt * 8 - 1);     // this pixel inside the band: granule zr/8, slot zr%8
    const int voff = (zr >> 3) * g.b_row + (zr & 7) * 2;
    float m = -INFINITY, l = 0.f, acc = 0.f;
    for (int T = 0; T < tiles; ++T) {
      const int s = T & 1;
      const int nv = min(G, n_img - T * G);
      mbar_wait(bar_pfull + 8 * s, (T >> 1) & 1, 4);
      mbar_wait(bar_tfull + 8 * s, (T >> 1) & 1, 5);
      mbar_wait(bar_bfull + 8 * s, (T >> 1) & 1, 6);        // already complete: acquires the loaders' shared-memory writes
      tc_fence_after();
      if (tid == 0) LS_STAMP(6, T);
      const uint8_t* bc = sBuf + (size_t)s * g.buf_bytes;
      const float* p2 = reinterpret_cast<const float*>(bc + g.b_bytes) + yl;
      const float* lws = reinterpret_cast<const float*>(bc + g.b_bytes + g.p2_bytes);
      for (int c0 = 16 * wg; c0 < G; c0 += 32) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tmem_base + lane_addr + s * G + c0, v);
        tmem_ld_wait16(v);
        if (c0 >= nv || (p.flags & 1)) continue;            // warp-uniform
        float t[16], lw[16], cmax = -INFINITY;
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(lw + e) = *reinterpret_cast<const float4*>(lws + c0 + e);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          t[e] = fmaf(__uint_as_float(v[e]), cs, fmaf(p2[(c0 + e) * 128], cn, lw[e]));   // lw = -inf masks a missing image
          cmax = fmaxf(cmax, t[e]);
        }
        if (cmax > m) {
          const float sc = ex2(m - cmax);
          l *= sc;
          acc *= sc;
          m = cmax;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float w = ex2(t[e] - m);
          l += w;
          acc = fmaf(w, __half2float(*reinterpret_cast<const __half*>(bc + voff + (c0 + e) * 16)), acc);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_done + 8 * s);
      if (tid == 0) LS_STAMP(7, T);
    }
    // merge the two groups' states (the raw ring is idle by now: every tile has been transposed and contracted)
    float* sMerge = reinterpret_cast<float*>(sRaw);
    if (wg == 1) {
      sMerge[yl * 3] = m;
      sMerge[yl * 3 + 1] = l;
      sMerge[yl * 3 + 2] = acc;
    }
    bar_sync_named(2, 256);
    if (wg == 0 && y < HW) {
      const float m1 = sMerge[yl * 3], l1 = sMerge[yl * 3 + 1], a1 = sMerge[yl * 3 + 2];
      const float M = fmaxf(m, m1);
      if (M > -INFINITY) {
        const float w0 = ex2(m - M), w1 = ex2(m1 - M);        // ex2(-inf) = 0: an empty side drops out
        l = fmaf(l, w0, l1 * w1);
        acc = fmaf(acc, w0, a1 * w1);
      }
      const size_t o = ((size_t)split * p.B + b) * HW + y;
      p.m[o] = M;
      p.l[o] = l;
      p.acc[o] = acc / p.scale;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void flat16_kernel(const float* __restrict__ images, long long N, int HW, int HWp, float scale, __half* __restrict__ out) {
  const long long total = N * HWp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int z = i % HWp;
    const long long n = i / HWp;
    out[i] = __float2half_rn(z < HW ? images[n * HW + z] * scale : 0.f);
  }
}

// window sums of T^2 (zero outside the image), fp32, [n][HWn], HWn = 128 * ceil(HW / 128)
__global__ void ls_norms_kernel(const float* __restrict__ images, long long N, int C, int H, int W, int k, int HWp,
                                float* __restrict__ out) {
  const int HW = H * W, d = k / 2;
  const long long total = N * HWp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = i % HWp;
    const long long n = i / HWp;
    float s2 = 0.f;
    if (y < HW) {
      const int ry = y / W, cy = y % W;
      for (int c = 0; c < C; ++c) {
        const float* img = images + (n * C + c) * HW;
        for (int r = max(0, ry - d); r <= min(H - 1, ry + d); ++r) {
          float rs = 0.f;
          for (int q = max(0, cy - d); q <= min(W - 1, cy + d); ++q) rs = fmaf(img[r * W + q], img[r * W + q], rs);
          s2 += rs;
        }
      }
    }
    out[i] = s2;
  }
}

}  // namespace

#ifdef CDS_PROFILE_SWITCHES
// profile builds only: timestamps of the first 64 tiles of CTA (0,0) (rows: landed, published, buffer free, issued, MMA start,
// MMA issued, epilogue start, epilogue end)
extern "C" int cds_debug_ls_clocks(long long* out_host) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out_host, g_ls_clk, sizeof(g_ls_clk));
  return e == cudaSuccess ? CDS_OK : CDS_ERR_CUDA;
}
#endif

extern "C" int64_t cds_ls_umma_smem_bytes(int C, int H, int W, int k, int passes) {
  LsUmmaGeom g;
  return make_ls_geom(C, H, W, k, passes, g) ? g.smem_total : 0;
}

extern "C" int64_t cds_ls_plane_elems(int64_t N, int H, int W) { return N * (((int64_t)H * W + 7) / 8 + 1) * 8; }

extern "C" int cds_pack_flat16(const float* images, int64_t N, int C, int H, int W, float scale, void* out_f16, void* stream) {
  CDS_CHECK_ARG(N >= 1 && C == 1 && H >= 1 && W >= 1, "cds_pack_flat16: single-channel images only");
  const int HW = H * W, HWp = ((HW + 7) / 8 + 1) * 8;
  const long long total = (long long)N * HWp;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  flat16_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, HW, HWp, scale, (__half*)out_f16);
  CDS_CHECK_LAUNCH("flat16_kernel");
  return CDS_OK;
}

extern "C" int64_t cds_ls_norms_elems(int64_t N, int H, int W) { return N * (((int64_t)H * W + 127) / 128) * 128; }

extern "C" int cds_pack_ls_norms(const float* images, int64_t N, int C, int H, int W, int k, float* out, void* stream) {
  CDS_CHECK_ARG(N >= 1 && C >= 1 && (k & 1) && k >= 1, "cds_pack_ls_norms: bad arguments");
  const int HWp = (H * W + 127) / 128 * 128;
  const long long total = (long long)N * HWp;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  ls_norms_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, C, H, W, k, HWp, out);
  CDS_CHECK_LAUNCH("ls_norms_kernel");
  return CDS_OK;
}

extern "C" int cds_ls_partials_umma(const float* x, int B, int C, int H, int W, int k, const float* beta, const void* flat16,
                                    float scale, const float* ls_norms, const int32_t* idx, const float* logw, int64_t n_sel,
                                    int splits, int passes, float* m, float* l, float* acc, void* stream) {
  LsUmmaParams p;
  if (!make_ls_geom(C, H, W, k, passes, p.g)) {
    cds_set_error("cds_ls_partials_umma: unsupported geometry C=%d H=%d W=%d k=%d passes=%d", C, H, W, k, passes);
    return CDS_ERR_UNSUPPORTED;
  }
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1 && scale > 0.f, "cds_ls_partials_umma: empty problem");
  if (splits > n_sel) splits = (int)n_sel;
  p.B = B; p.splits = splits; p.n_sel = n_sel;
  p.x = x; p.beta = beta;
  p.plane = (const __half*)flat16; p.norms = ls_norms;
  p.scale = scale; p.idx = idx; p.logw = logw;
  p.m = m; p.l = l; p.acc = acc;
  {
    const char* f = getenv("CDS_LS_DEBUG");
    p.flags = f ? atoi(f) : 0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(ls_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g.smem_total);
  if (e != cudaSuccess) {
    cds_set_error("cds_ls_partials_umma: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  ls_umma_kernel<<<dim3(p.g.MT * B, splits), LS_THREADS, p.g.smem_total, st>>>(p);
  CDS_CHECK_LAUNCH("ls_umma_kernel");
  return CDS_OK;
}
