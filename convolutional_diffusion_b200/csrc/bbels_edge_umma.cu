// bbELS edge bands on the tensor cores (tcgen05 UMMA, accumulators in TMEM).
//
// Reference behaviour restated (never copied): /root/reference/src/utils/idealscore.py:256-288.  A query pixel in an
// edge band at depth r from its border is compared with the zero-padded patches of every bank image centred at the SAME
// depth r and at every interior position along the band (see csrc/bbels_edge.cu for the exact fp32 SIMT form).
//
// Why this is one GEMM per band and not one per depth: bring the band to one orientation (view[t][s], t = across the band
// starting at the border, s = along it).  The patch of the candidate at (depth r, position j) covers the view rows
// 0 .. r+d and the columns j-d .. j+d; the rows above the border are zero padding on the query and the candidate alike.
// So  q(r,i) . p(r,j) = sum over ALL view rows t < k-1 of  qz_r[t][.] * view[t][.]  where qz_r is the query patch with
// its rows t > r+d zeroed: the depth truncation lives entirely on the query operand, and the candidate operand -- the
// k-1 view rows under the border, as 8-pixel granules across the band -- does not depend on the depth at all.  All
// depths of a band therefore stack on the M side of ONE contraction
//     S[(r,i), (image, j)] = A[(r,i), K] . B[K, (image, j)],   K = C * (k-1 rows as granules of 8) * k columns,
// with the same strip aliasing as the ELS kernel on both operands (granule (c, g, s) = 8 view rows 8g..8g+7 at along
// position s; the K slice "dx" of candidate j is the granule at s = j-d+dx, so 8 consecutive candidates are one 128-byte
// core matrix at a 16-byte offset per dx).  The only depth-dependent candidate data are the truncated-patch norms
// |p(r,j)|^2 -- one more K granule per depth, multiplied by -a*scale/2 in the query rows of that depth and by 0 in the
// others -- and the value (centre pixel = view[r][j]), which the epilogue reads from the staged granules.
//
// M rows: group (h, r) of 8 consecutive positions i = 8h..8h+7 at depth r, ordered h-major so that the distance between
// groups is uniform (UMMA SBO): the query block of a (c, g) is [NH*d groups][k+7 granules], the rows of h > 0 are copies
// shifted by 8 granules.  k=17 on 32x32: 2*8 groups = 128 rows, one M tile, 55 K steps, N = 8*G candidates per
// instruction (G images, one instruction per 8-position group h').
//
// CTA = (band, M tile, bank slice, sample); warps 0-3 epilogue (TMEM lane = query row), warps 4-7 producers (bulk copies of
// the per-image granule rows + norm granules into a 2-stage ring), warp 8 MMA issuer (two TMEM accumulator buffers).
// passes = 1: fp16 query; 2: fp16 hi + lo residual of the query (the K steps of the strips run twice), as in the ELS kernel.
// The bank must be a single exact fp16 plane; everything else stays on the SIMT kernel.
#include "umma_common.cuh"

namespace {
using namespace umma;

constexpr int EDGE_THREADS = 288;     // warps 0-3 epilogue, 4-7 producers (one per copy kind), 8 MMA issuer
constexpr int EDGE_MAX_MMAS = 400;

struct EdgeGeom {
  int C, H, k, D, I, passes;
  int NH;            // 8-position groups along the band = ceil(I/8)
  int NGR;           // 8-row granules across the band that a patch can touch = ceil((k-1)/8)
  int NGRall;        // granule rows stored per channel in the edge plane = ceil(H/8)
  int MT;            // M tiles = ceil(NH*D/16)
  int RA, a_block;   // query row-group stride (k+7 granules), bytes of one (c,g) query block (16 groups)
  int a_plane;        // bytes of one precision plane of the query blocks = C*NGR*a_block
  int a_zero, a_norm, a_nzero, a_bytes;
  int S1;            // bytes of one granule row in a stage = H*16
  int strip_bytes;   // C*NGR*S1
  int norm_bytes;    // D*NH*128
  int img_stride;    // bytes per image in a stage
  int G;             // images per tile
  int ncol;          // TMEM columns per accumulator buffer = NH*8*G
  int n_strip, n_mma;
  int stage_bytes, smem_A, smem_total;
};

struct EdgeUmmaParams {
  EdgeGeom g;
  int B, splits;
  long long n_sel;
  const float* x;
  const float* beta;
  const uint8_t* plane;      // [n][band][c][NGRall][H] granules
  const uint8_t* norms;      // [n][band][D][8*NH] granules
  float scale;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc;
  uint2 table[EDGE_MAX_MMAS];
};

// image coordinates of view element (t, s): band 0 top, 1 bottom, 2 left, 3 right (square images)
__host__ __device__ __forceinline__ int view_pixel(int band, int t, int s, int H) {
  switch (band) {
    case 0: return t * H + s;
    case 1: return (H - 1 - t) * H + s;
    case 2: return s * H + t;
    default: return s * H + (H - 1 - t);
  }
}

bool make_edge_geom(int C, int H, int k, int passes, EdgeGeom& g, uint2* table) {
  if (passes < 1 || passes > 2) return false;
  if (C < 1 || C > 3 || (k & 1) == 0 || k < 3 || k >= H || H > 64 || H < 8) return false;
  g.C = C; g.H = H; g.k = k; g.D = k / 2; g.I = H - 2 * g.D; g.passes = passes;
  if (g.I < 1) return false;
  g.NH = (g.I + 7) / 8;
  g.NGR = (k - 1 + 7) / 8;
  g.NGRall = (H + 7) / 8;
  g.MT = (g.NH * g.D + 15) / 16;
  g.RA = (k + 7) * 16;
  g.a_block = 16 * g.RA;
  g.a_plane = C * g.NGR * g.a_block;
  g.a_zero = passes * g.a_plane;
  g.a_norm = g.a_zero + g.a_block;
  g.a_nzero = g.a_norm + g.D * 2048;
  g.a_bytes = g.a_nzero + 2048;
  g.S1 = H * 16;
  g.strip_bytes = C * g.NGR * g.S1;
  g.norm_bytes = g.D * g.NH * 128;
  g.img_stride = g.strip_bytes + g.norm_bytes;
  g.smem_A = (g.a_bytes + 1023) / 1024 * 1024;
  g.G = 0;
  for (int G = 16; G >= 2; G >>= 1) {
    if (g.NH * 8 * G > 256) continue;
    if (g.smem_A + 2 * G * g.img_stride + 1024 + 128 <= 227 * 1024) { g.G = G; break; }
  }
  if (g.G == 0) return false;
  g.ncol = g.NH * 8 * g.G;
  g.stage_bytes = g.G * g.img_stride;
  g.smem_total = g.smem_A + 2 * g.stage_bytes + 1024 + 128;      // 1 KB guard: masked columns read past the last image
  if ((g.img_stride >> 4) > 0x3FFF) return false;
  // K granule lists (offsets relative to the query tile / to the image's slot in a stage)
  static thread_local Gran gr[3 * 8 * 64 + 2];
  int n = 0;
  for (int c = 0; c < C; ++c)
    for (int gg = 0; gg < g.NGR; ++gg)
      for (int dx = 0; dx < k; ++dx) {
        gr[n].a = (c * g.NGR + gg) * g.a_block + dx * 16;
        gr[n].b = (c * g.NGR + gg) * g.S1 + dx * 16;
        ++n;
      }
  if (passes * ((n + 1) / 2) + (g.D + 1) / 2 > EDGE_MAX_MMAS) return false;
  int nm = 0;
  for (int pa = 0; pa < passes; ++pa) {   // pass 1: the fp16 residual of the query against the same candidate granules
    for (int q = 0; q < n; q += 2) {
      const int a0 = gr[q].a + pa * g.a_plane;
      int la, lb;
      if (q + 1 < n) { la = gr[q + 1].a - gr[q].a; lb = gr[q + 1].b - gr[q].b; }
      else           { la = g.a_zero - a0;         lb = 16; }
      if (la <= 0 || lb <= 0 || (la >> 4) > 0x3FFF || (lb >> 4) > 0x3FFF) return false;
      table[nm].x = (uint32_t)(a0 >> 4) | ((uint32_t)(la >> 4) << 16);
      table[nm].y = (uint32_t)(gr[q].b >> 4) | ((uint32_t)(lb >> 4) << 16);
      ++nm;
    }
  }
  g.n_strip = nm;
  for (int d = 0; d < g.D; d += 2) {
    const int a0 = g.a_norm + d * 2048, b0 = g.strip_bytes + d * g.NH * 128;
    const int la = d + 1 < g.D ? 2048 : g.a_nzero - a0, lb = d + 1 < g.D ? g.NH * 128 : 16;
    table[nm].x = (uint32_t)(a0 >> 4) | ((uint32_t)(la >> 4) << 16);
    table[nm].y = (uint32_t)(b0 >> 4) | ((uint32_t)(lb >> 4) << 16);
    ++nm;
  }
  g.n_mma = nm;
  return true;
}

template <int C>
__global__ void __launch_bounds__(EDGE_THREADS, 1) bbels_edge_umma_kernel(const __grid_constant__ EdgeUmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const EdgeGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int band = blockIdx.x / g.MT, mt = blockIdx.x % g.MT;
  const int split = blockIdx.y, b = blockIdx.z;
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  const int n_img = (int)(n1 - n0);
  const int tiles = (n_img + g.G - 1) / g.G;
  const int H = g.H, D = g.D, I = g.I, HW = H * H;

  uint8_t* sA = smem;
  uint8_t* sStage = smem + g.smem_A;
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + g.smem_A + 2 * g.stage_bytes + 1024);
  // barriers: full[2] (bulk copies landed), tfull[2] (accumulator tile complete), done[2] (epilogue finished with the stage
  // and the accumulator buffer of the same index)
  const uint32_t bar_full = smem_u32(sBar), bar_tfull = bar_full + 16, bar_done = bar_full + 32;
  uint32_t* sTmemBase = reinterpret_cast<uint32_t*>(sBar + 6);

  const float beta = p.beta[b];
  const float a = sqrtf(1.f - beta);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, C + 1);    // one producer warp per channel plane + one for the norms (expect_tx each)
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_done + 8 * s, 4);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(sTmemBase), TMEM_COLS);
  // masked candidate columns and zero-weighted K slots read whatever lies in the ring: keep it finite
  for (int e = tid * 16; e < 2 * g.stage_bytes + 1024; e += EDGE_THREADS * 16)
    *reinterpret_cast<uint4*>(sStage + e) = make_uint4(0, 0, 0, 0);
  {
    // query blocks: block (c,gg), row group gl <-> (h, r) of this M tile, granule s' = the 8 view rows 8gg..8gg+7 of x at along
    // position 8h+s', rows deeper than r+d (outside the depth-r patch) and positions past the image zeroed
    const float* xb = p.x + (size_t)b * C * HW;
    const int JW = g.k + 7;
    const int per = C * g.NGR * 16 * JW;
    for (int e = tid; e < per; e += EDGE_THREADS) {
      const int sp = e % JW, gl = (e / JW) & 15, blk = e / (JW * 16);
      const int c = blk / g.NGR, gg = blk % g.NGR;
      const int gidx = mt * 16 + gl, h = gidx / D, r = gidx % D;
      const int s = 8 * h + sp;
      __half v[8], vl[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int t = 8 * gg + q;
        const bool ok = h < g.NH && s < H && t <= r + D && t < H;
        const float xv = ok ? xb[c * HW + view_pixel(band, t, s, H)] : 0.f;
        v[q] = __float2half_rn(xv);
        vl[q] = __float2half_rn(xv - __half2float(v[q]));
      }
      const size_t off = (size_t)blk * g.a_block + (size_t)gl * g.RA + (size_t)sp * 16;
      *reinterpret_cast<uint4*>(sA + off) = *reinterpret_cast<uint4*>(v);
      if (g.passes > 1) *reinterpret_cast<uint4*>(sA + g.a_plane + off) = *reinterpret_cast<uint4*>(vl);
    }
    for (int e = tid * 16; e < g.a_block; e += EDGE_THREADS * 16)
      *reinterpret_cast<uint4*>(sA + g.a_zero + e) = make_uint4(0, 0, 0, 0);
    // norm coefficient granules: depth d's K granule is (gh,gh,gh,gm,gm,gl,0,0) (against the plane's (ph,pm,pl,ph,pm,ph,0,0))
    // in the rows of depth d and zero in all others
    const float gamma = -0.5f * a * p.scale;
    const __half gh = __float2half_rn(gamma);
    const __half gm = __float2half_rn(gamma - __half2float(gh));
    const __half gl3 = __float2half_rn(gamma - __half2float(gh) - __half2float(gm));
    const __half z = __float2half_rn(0.f);
    __half cg[8] = {gh, gh, gh, gm, gm, gl3, z, z};
    const uint4 cgv = *reinterpret_cast<uint4*>(cg);
    for (int e = tid; e < (D + 1) * 128; e += EDGE_THREADS) {       // D blocks of 16 groups x 8 rows, then the zero block
      const int d = e >> 7, gl = (e >> 3) & 15;
      const int gidx = mt * 16 + gl, h = gidx / D, r = gidx % D;
      const bool on = d < D && h < g.NH && r == d;
      *reinterpret_cast<uint4*>(sA + g.a_norm + (size_t)e * 16) = on ? cgv : make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmemBase;

  if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- producers.  A warp gets a cp.async.bulk out every ~100
    // cycles (measured in the LS kernel), so the (C+1) copies per image go to C+1 warps: warp 4+c the granule rows of channel c,
    // warp 4+C the norm granules; one copy per lane and tile (G <= 16 images), index lookups one tile ahead.
    const int part = warp - 4;
    if (part <= C) {
      const size_t img_bytes = (size_t)4 * C * g.NGRall * g.S1, band_bytes = (size_t)C * g.NGRall * g.S1;
      const size_t nimg_bytes = (size_t)4 * g.norm_bytes;
      const int per_c = g.NGR * g.S1;
      const uint32_t bytes = part < C ? (uint32_t)per_c : (uint32_t)g.norm_bytes;
      const uint8_t* src0 = part < C ? p.plane + band * band_bytes + (size_t)part * g.NGRall * g.S1 : p.norms + (size_t)band * g.norm_bytes;
      const size_t stride = part < C ? img_bytes : nimg_bytes;
      const uint32_t off = part < C ? (uint32_t)(part * per_c) : (uint32_t)g.strip_bytes;
      auto lookup = [&](int T) -> long long {
        return (T < tiles && lane < min(g.G, n_img - T * g.G)) ? (long long)p.idx[n0 + (long long)T * g.G + lane] : -1;
      };
      long long img = lookup(0);
      for (int T = 0; T < tiles; ++T) {
        const int s = T & 1;
        const long long next = lookup(T + 1);
        mbar_wait(bar_done + 8 * s, ((T >> 1) & 1) ^ 1, 1);
        const int nv = min(g.G, n_img - T * g.G);
        if (lane == 0) mbar_expect_tx(bar_full + 8 * s, (uint32_t)nv * bytes);
        __syncwarp();
        if (img >= 0)
          bulk_g2s(smem_u32(sStage + (size_t)s * g.stage_bytes) + lane * g.img_stride + off, src0 + (size_t)img * stride, bytes,
                   bar_full + 8 * s);
        img = next;
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------------------------- MMA issuer
    const uint64_t a_hi = desc_hi(g.RA), a_hi_n = desc_hi(128), b_hi = desc_hi(g.img_stride);
    const uint32_t idesc = (1u << 4) | (((uint32_t)(8 * g.G) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(sA) >> 4;
    for (int T = 0; T < tiles; ++T) {
      const int s = T & 1;
      // full[s] of tile T implies done[s] of tile T-2: the producer refills a stage only after the epilogue released it,
      // and the accumulator buffer of the same index is released by the same arrival
      mbar_wait(bar_full + 8 * s, (T >> 1) & 1, 2);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_base = smem_u32(sStage + (size_t)s * g.stage_bytes) >> 4;
        for (int hc = 0; hc < g.NH; ++hc) {
          const uint32_t d_tmem = tmem_base + s * 256 + hc * 8 * g.G;
          const uint32_t bb = b_base + hc * 8;
          for (int t = 0; t < g.n_mma; ++t) {
            const uint2 e = p.table[t];
            umma_f16(d_tmem, (t < g.n_strip ? a_hi : a_hi_n) | (uint64_t)(e.x + a_base), b_hi | (uint64_t)(e.y + bb), idesc,
                     t ? 1u : 0u);
          }
        }
        umma_commit(bar_tfull + 8 * s);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = query row
    const int q = tid;
    const int gidx = mt * 16 + (q >> 3), hq = gidx / D, r = gidx % D, iq = 8 * hq + (q & 7);
    const bool qvalid = hq < g.NH && iq < I;
    const float cs = a * CDS_LOG2E / (beta * p.scale);
    const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;
    // value of candidate (depth r, position i): granule (c, r/8, s = d+i), slot r%8
    const int voff = ((r >> 3) * H + D) * 16 + (r & 7) * 2, vstep = g.NGR * g.S1;
    float m = -INFINITY, l = 0.f, acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int T = 0; T < tiles; ++T) {
      const int s = T & 1;
      const int nv = min(g.G, n_img - T * g.G);
      const float lw_mine = lane < nv ? __ldg(p.logw + n0 + (long long)T * g.G + lane) * CDS_LOG2E : -INFINITY;
      mbar_wait(bar_tfull + 8 * s, (T >> 1) & 1, 3);
      mbar_wait(bar_full + 8 * s, (T >> 1) & 1, 4);     // already complete: acquires the bulk-copied granules for the value reads
      tc_fence_after();
      const uint8_t* st = sStage + (size_t)s * g.stage_bytes;
      for (int hc = 0; hc < g.NH; ++hc) {
        const int ncand = min(8, I - 8 * hc);
        for (int i2 = 0; i2 < g.G / 2; ++i2) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(tmem_base + lane_addr + s * 256 + hc * 8 * g.G + i2 * 16, v);
          tmem_ld_wait16(v);
          const float lw0 = __shfl_sync(0xffffffffu, lw_mine, 2 * i2), lw1 = __shfl_sync(0xffffffffu, lw_mine, 2 * i2 + 1);
          if (2 * i2 >= nv) continue;           // warp-uniform: both images beyond the slice
          float t[16], cmax = -INFINITY;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const bool ok = (e & 7) < ncand;
            t[e] = ok ? fmaf(__uint_as_float(v[e]), cs, e < 8 ? lw0 : lw1) : -INFINITY;      // lw = -inf masks a missing image
            cmax = fmaxf(cmax, t[e]);
          }
          if (cmax == -INFINITY) continue;
          if (cmax > m) {
            const float sc = ex2(m - cmax);
            l *= sc;
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] *= sc;
            m = cmax;
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float w = ex2(t[e] - m);
            l += w;
            const uint8_t* vp = st + (size_t)(2 * i2 + (e >> 3)) * g.img_stride + voff + (8 * hc + (e & 7)) * 16;
#pragma unroll
            for (int c = 0; c < C; ++c)
              acc[c] = fmaf(w, __half2float(*reinterpret_cast<const __half*>(vp + c * vstep)), acc[c]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_done + 8 * s);
    }
    if (qvalid) {
      const int pix = view_pixel(band, r, D + iq, H);
      const size_t o = ((size_t)split * p.B + b) * HW + pix;
      const float inv = 1.f / p.scale;
      p.m[o] = m;
      p.l[o] = l;
#pragma unroll
      for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b) * C + c) * HW + pix] = acc[c] * inv;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// edge plane: granule (n, band, c, g, s) = 8 view rows 8g..8g+7 at along position s, fp16(pixel * scale), zero past the image
__global__ void edge_plane_kernel(const float* __restrict__ images, long long N, int C, int H, float scale,
                                  uint4* __restrict__ out) {
  const int NG = (H + 7) / 8;
  const long long total = N * 4 * C * NG * H;
  for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
    const int s = gi % H, gg = (gi / H) % NG, c = (gi / ((long long)H * NG)) % C, band = (gi / ((long long)H * NG * C)) & 3;
    const long long n = gi / ((long long)H * NG * C * 4);
    const float* src = images + (n * C + c) * H * H;
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int t = 8 * gg + e;
      h[e] = __float2half_rn(t < H ? src[view_pixel(band, t, s, H)] * scale : 0.f);
    }
    out[gi] = *reinterpret_cast<uint4*>(h);
  }
}

// edge norms: granule (n, band, d, i) = three-way fp16 split of the squared norm of the zero-padded patch centred at depth d,
// interior position i (view rows 0..d+D, columns i..i+k-1), arranged (ph,pm,pl,ph,pm,ph,0,0); zero for i >= I
__global__ void edge_norms_kernel(const float* __restrict__ images, long long N, int C, int H, int k, uint4* __restrict__ out) {
  const int D = k / 2, I = H - 2 * D, NHP = (I + 7) / 8 * 8;
  const long long total = N * 4 * D * NHP;
  for (long long gi = blockIdx.x * (long long)blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
    const int i = gi % NHP, d = (gi / NHP) % D, band = (gi / ((long long)NHP * D)) & 3;
    const long long n = gi / ((long long)NHP * D * 4);
    __half h[8];
    const __half z = __float2half_rn(0.f);
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = z;
    if (i < I) {
      const float* img = images + n * C * H * H;
      float s2 = 0.f;
      for (int c = 0; c < C; ++c)
        for (int t = 0; t <= d + D; ++t) {
          float rs = 0.f;
          for (int q = 0; q < k; ++q) {
            const float v = img[c * H * H + view_pixel(band, t, i + q, H)];
            rs = fmaf(v, v, rs);
          }
          s2 += rs;
        }
      const __half ph = __float2half_rn(s2);
      const __half pm = __float2half_rn(s2 - __half2float(ph));
      const __half pl = __float2half_rn(s2 - __half2float(ph) - __half2float(pm));
      h[0] = ph; h[1] = pm; h[2] = pl; h[3] = ph; h[4] = pm; h[5] = ph;
    }
    out[gi] = *reinterpret_cast<uint4*>(h);
  }
}

}  // namespace

extern "C" int64_t cds_bbels_edge_umma_smem_bytes(int C, int H, int W, int k, int passes) {
  if (H != W) return 0;
  EdgeGeom g;
  static thread_local uint2 scratch[EDGE_MAX_MMAS];
  return make_edge_geom(C, H, k, passes, g, scratch) ? g.smem_total : 0;
}

extern "C" int64_t cds_edge_plane_halves(int64_t N, int C, int H) { return N * 4 * C * ((H + 7) / 8) * H * 8; }
extern "C" int64_t cds_edge_norms_halves(int64_t N, int H, int k) {
  const int D = k / 2, I = H - 2 * D;
  return I < 1 ? 0 : N * 4 * D * ((I + 7) / 8 * 8) * 8;
}

extern "C" int cds_pack_edge_plane(const float* images, int64_t N, int C, int H, float scale, void* out_f16, void* stream) {
  CDS_CHECK_ARG(N >= 1 && C >= 1 && H >= 1, "cds_pack_edge_plane: bad arguments");
  const long long total = (long long)N * 4 * C * ((H + 7) / 8) * H;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  edge_plane_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, C, H, scale, (uint4*)out_f16);
  CDS_CHECK_LAUNCH("edge_plane_kernel");
  return CDS_OK;
}

extern "C" int cds_pack_edge_norms(const float* images, int64_t N, int C, int H, int k, void* out_f16, void* stream) {
  CDS_CHECK_ARG(N >= 1 && (k & 1) && k >= 3 && k < H, "cds_pack_edge_norms: bad arguments");
  const int D = k / 2, I = H - 2 * D;
  const long long total = (long long)N * 4 * D * ((I + 7) / 8 * 8);
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  edge_norms_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, C, H, k, (uint4*)out_f16);
  CDS_CHECK_LAUNCH("edge_norms_kernel");
  return CDS_OK;
}

extern "C" int cds_bbels_edge_partials_umma(const float* x, int B, int C, int H, int W, int k, const float* beta,
                                            const void* edge_plane, float scale, const void* edge_norms, const int32_t* idx,
                                            const float* logw, int64_t n_sel, int splits, int passes, float* m, float* l,
                                            float* acc, void* stream) {
  static thread_local EdgeUmmaParams p;
  if (H != W || !make_edge_geom(C, H, k, passes, p.g, p.table)) {
    cds_set_error("cds_bbels_edge_partials_umma: unsupported geometry C=%d H=%d W=%d k=%d passes=%d", C, H, W, k, passes);
    return CDS_ERR_UNSUPPORTED;
  }
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1 && scale > 0.f, "cds_bbels_edge_partials_umma: empty problem");
  if (splits > n_sel) splits = (int)n_sel;
  p.B = B; p.splits = splits; p.n_sel = n_sel;
  p.x = x; p.beta = beta;
  p.plane = (const uint8_t*)edge_plane; p.norms = (const uint8_t*)edge_norms;
  p.scale = scale; p.idx = idx; p.logw = logw;
  p.m = m; p.l = l; p.acc = acc;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(4 * p.g.MT, splits, B);
  cudaError_t e = cudaSuccess;
#define LAUNCH(CC)                                                                                                        \
  e = cudaFuncSetAttribute(bbels_edge_umma_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g.smem_total);      \
  if (e == cudaSuccess) bbels_edge_umma_kernel<CC><<<grid, EDGE_THREADS, p.g.smem_total, st>>>(p);
  if (C == 1) { LAUNCH(1) } else if (C == 2) { LAUNCH(2) } else { LAUNCH(3) }
#undef LAUNCH
  if (e != cudaSuccess) {
    cds_set_error("cds_bbels_edge_partials_umma: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  CDS_CHECK_LAUNCH("bbels_edge_umma_kernel");
  return CDS_OK;
}
