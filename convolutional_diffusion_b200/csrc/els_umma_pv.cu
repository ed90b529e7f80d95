// ELS score partials, "P.V" variant: the softmax-weighted sum of the centre pixels also runs on the tensor cores.
//
// Same contraction, staging and descriptor aliasing as els_umma.cu (see there).  What differs is the epilogue:
//   S tile (128 queries x N <= 240 candidates, fp32, TMEM) --epilogue--> P = 2^(logit - m_ref + 14) in fp16, written
//   back INTO the S buffer (tcgen05.st; P of a 16-column chunk occupies 8 columns at the start of its owner's column
//   range, so unread S columns are never clobbered) --second UMMA--> O[128 x 16] += P[128 x N] . V'[N x 16], with
//   V' = (scale*T_c ..., 1, residual planes ...) staged per tile as an fp16 K-major operand by the builder warps.
//   Column C of O is the softmax denominator.  One warpgroup folds the 16-column O tiles into the per-query
//   (m, l, acc) state; nothing else of the weighted sum touches the FMA pipe or shared memory.
// Per tile the four epilogue warpgroups agree on one reference max per query (shared memory + one named barrier),
// because P.V sums over all columns of the tile.
// TMEM: S buffers at columns 0 and 240, O tiles at 480 and 496.
#include "umma_common.cuh"

using namespace umma;

namespace {

constexpr int S_BUF_COLS = 240;
constexpr int O_COL0 = 480;
constexpr float P_SHIFT = 14.f;          // P is stored as 2^(logit - m_ref + 14) <= 16384 (fp16 keeps 2^-38 .. 2^14)

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait8(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <int C>
__global__ void __launch_bounds__(THREADS, 1) els_umma_pv_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const UmmaGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_j = (g.W + TJ - 1) / TJ;
  const int i0 = (blockIdx.x / tiles_j) * TI, j0 = (blockIdx.x % tiles_j) * TJ;
  const int split = blockIdx.y, b = blockIdx.z;
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  const int n_img = (int)(n1 - n0);
  const int S = g.stages;
  const int tiles_per_img = g.nchunks * g.nvb;
  const int total_tiles = n_img * tiles_per_img;

  uint8_t* sA = smem;
  uint8_t* sStage = smem + g.smem_A;
  float* sMax = reinterpret_cast<float*>(smem + g.smem_A + g.smem_stage);   // [2][NUM_EPI_WG][128] per-tile chunk maxima
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + g.smem_A + g.smem_stage + g.smem_merge + g.smem_table);
  // barriers: full[2], empty[2], vready[2], tfull[2], pready[2], oready[2]; then the TMEM base address
  const uint32_t bar_full = smem_u32(sBar), bar_empty = bar_full + 16, bar_vready = bar_full + 32;
  const uint32_t bar_tfull = bar_full + 48, bar_pready = bar_full + 64, bar_oready = bar_full + 80;
  uint32_t* sTmemBase = reinterpret_cast<uint32_t*>(sBar + 12);

  const float beta = p.beta[b];
  const float a = sqrtf(1.f - beta);
  const float inv_scale = 1.f / p.scale;

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);                    // MMA commit after the last P.V of the image
      mbar_init(bar_vready + 8 * s, 2);                   // the two builder warps
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(bar_tfull + 8 * q, 1);
      mbar_init(bar_pready + 8 * q, 4 * NUM_EPI_WG);      // every epilogue warp has written its share of P
      mbar_init(bar_oready + 8 * q, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(sTmemBase), TMEM_COLS);
  for (int e = tid * 16; e < g.smem_stage; e += THREADS * 16) *reinterpret_cast<uint4*>(sStage + e) = make_uint4(0, 0, 0, 0);
  {
    // query tile (A operand), constant block and zero block: identical to els_umma.cu
    const int JW = g.k + 7;
    const int per_plane = C * g.nb * TI * JW;
    const float* xb = p.x + (size_t)b * C * g.H * g.W;
    for (int e = tid; e < per_plane; e += THREADS) {
      const int jj = e % JW, gi = (e / JW) % TI, blk = (e / (JW * TI)) % g.nb, c = e / (JW * TI * g.nb);
      __half hi[8], lo[8];
      int xc = j0 + jj - g.d;
      bool colok = true;
      if (p.pad == CDS_PAD_CIRCULAR) xc = ((xc % g.W) + g.W) % g.W;
      else colok = (xc >= 0 && xc < g.W);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int dy = 8 * blk + q;
        int yr = i0 + gi + dy - g.d;
        bool ok = colok && dy < g.k;
        if (p.pad == CDS_PAD_CIRCULAR) yr = ((yr % g.H) + g.H) % g.H;
        else ok = ok && (yr >= 0 && yr < g.H);
        const float v = ok ? xb[(c * g.H + yr) * g.W + xc] : 0.f;
        hi[q] = __float2half_rn(v);
        lo[q] = __float2half_rn(v - __half2float(hi[q]));
      }
      const size_t off = (size_t)(c * g.nb + blk) * g.a_block + (size_t)gi * g.RA + (size_t)jj * 16;
      *reinterpret_cast<uint4*>(sA + off) = *reinterpret_cast<uint4*>(hi);
      if (g.passes > 1) *reinterpret_cast<uint4*>(sA + g.a_plane + off) = *reinterpret_cast<uint4*>(lo);
    }
    const float gamma = -0.5f * a * p.scale;
    const __half gh = __float2half_rn(gamma);
    const __half gm = __float2half_rn(gamma - __half2float(gh));
    const __half gl = __float2half_rn(gamma - __half2float(gh) - __half2float(gm));
    const __half z = __float2half_rn(0.f);
    __half cg[8] = {gh, gh, gh, gm, gm, gl, z, z};
    for (int e = tid * 16; e < g.a_block; e += THREADS * 16) {
      *reinterpret_cast<uint4*>(sA + g.a_const + e) = *reinterpret_cast<uint4*>(cg);
      *reinterpret_cast<uint4*>(sA + g.a_zero + e) = make_uint4(0, 0, 0, 0);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sTmemBase;

  if (warp == 0) {
    // =========================== producer (one band per image in this variant: unit == image)
    if (lane == 0) {
      const uint32_t rows = (uint32_t)min(g.R, g.H), nrows = (uint32_t)min(g.G, g.H);
      const uint32_t cbytes = rows * g.S1, nbytes = nrows * g.S1;
      for (int n = 0; n < n_img; ++n) {
        const int s = n % S;
        mbar_wait(bar_empty + 8 * s, ((n / S) & 1) ^ 1, 1);
        const long long gi = p.idx[n0 + n];
        const uint32_t dst = smem_u32(sStage + (size_t)s * g.stage_bytes);
        mbar_expect_tx(bar_full + 8 * s, g.bank_planes * g.C * cbytes + nbytes);
        for (int c = 0; c < g.C; ++c) {
          const size_t src = ((size_t)gi * g.C + c) * g.chan_bytes;
          bulk_g2s(dst + c * g.R * g.S1, p.bank_hi + src, cbytes, bar_full + 8 * s);
          if (g.bank_planes > 1)
            bulk_g2s(dst + g.img_bytes + g.tile_pad + c * g.R * g.S1, p.bank_lo + src, cbytes, bar_full + 8 * s);
        }
        bulk_g2s(dst + g.np_off, p.norm_plane + (size_t)gi * g.chan_bytes, nbytes, bar_full + 8 * s);
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer.  Issue order: S(0), S(1), P.V(0), S(2), P.V(1), S(3), ...  S(T+2) reuses
    // the buffer of tile T and is issued after P.V(T), which itself waits until every epilogue warp has finished
    // reading S(T) and writing P(T); tensor-core operations execute in issue order.
    const uint64_t a_hi = desc_hi(g.RA), b_hi = desc_hi(g.S1);
    const uint64_t v_hi = desc_hi(128);                 // V' operand: 8-row groups 128 B apart, K granules 256 B apart
    const uint32_t a_base = smem_u32(sA) >> 4;
    const int nm = g.n_mma;
    // P.V instruction: D=f32, A/B=f16, K-major, N=16, M=128
    const uint32_t idesc_pv = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

    auto issue_pv = [&](int T) {
      // tile T of image T / tiles_per_img: chunk / column-block indices, operands, barriers
      const int n = T / tiles_per_img, tt = T - n * tiles_per_img;
      const int ch = tt / g.nvb;
      const int s = n % S;
      const int buf = T & 1;
      const int nck = (8 * g.chunk_g[ch]) >> 4;                         // 16-column chunks of this tile
      const int cw = (nck + NUM_EPI_WG - 1) / NUM_EPI_WG;               // chunks per warpgroup
      if (tt == 0) mbar_wait(bar_vready + 8 * s, (n / S) & 1, 7);      // V' operands of this image are built
      mbar_wait(bar_pready + 8 * buf, (uint32_t)((T >> 1) & 1), 8);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vt = (smem_u32(sStage + (size_t)s * g.stage_bytes + g.vt_off) + (uint32_t)tt * g.vt_tile * 4u) >> 4;
        const uint32_t d_o = tmem_base + O_COL0 + buf * 16;
        const uint32_t s_col = tmem_base + buf * S_BUF_COLS;
        for (int j = 0; j < nck; ++j) {
          const int w = j / cw, i = j - w * cw;
          const uint32_t a_t = s_col + 16 * w * cw + 8 * i;             // P of chunk j (see epilogue)
          const uint64_t bd = v_hi | (uint64_t)(((vt + (uint32_t)j * 32u) & 0x3FFFu) | (16u << 16));   // +512 B per chunk, LBO 256 B
          umma_f16_ts(d_o, a_t, bd, idesc_pv, j > 0 ? 1u : 0u);
        }
        umma_commit(bar_oready + 8 * buf);
        if (tt == tiles_per_img - 1) umma_commit(bar_empty + 8 * s);    // last reader of this stage
      }
      __syncwarp();
    };

    int T = 0;
    for (int n = 0; n < n_img; ++n) {
      const int s = n % S;
      const uint32_t stage_addr = smem_u32(sStage + (size_t)s * g.stage_bytes);
      for (int ch = 0; ch < g.nchunks; ++ch) {
        const uint32_t N = 8u * g.chunk_g[ch];
        const uint32_t idesc = (1u << 4) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        for (int vb = 0; vb < g.nvb; ++vb, ++T) {
          // the pending P.V first: it may be the commit that releases the stage the next image is waiting for
          if (T >= 2) issue_pv(T - 2);
          if (ch == 0 && vb == 0) {
            mbar_wait(bar_full + 8 * s, (n / S) & 1, 2);
            tc_fence_after();
          }
          const int buf = T & 1;
          const uint32_t d_tmem = tmem_base + buf * S_BUF_COLS;
          const uint32_t b_base = (stage_addr + vb * 128u) >> 4;
          if (elect_one()) {
            {
              const uint2 e = p.table[0];
              umma_f16(d_tmem, a_hi | (uint64_t)(e.x + a_base), b_hi | (uint64_t)(e.y + b_base), idesc, 0u);
            }
#pragma unroll 4
            for (int t = 1; t < nm; ++t) {
              const uint2 e = p.table[t];
              umma_f16(d_tmem, a_hi | (uint64_t)(e.x + a_base), b_hi | (uint64_t)(e.y + b_base), idesc, 1u);
            }
            umma_commit(bar_tfull + 8 * buf);
          }
          __syncwarp();
        }
      }
    }
    if (total_tiles >= 2) issue_pv(total_tiles - 2);
    issue_pv(total_tiles - 1);
  } else if (warp == 2 || warp == 3) {
    // =========================== builders: V'^T operand of every tile of the staged image, fp16, K-major no-swizzle:
    // element (row nrow, candidate column r) at  (r/8)*256 + (nrow/8)*128 + (nrow%8)*16 + (r%8)*2  bytes;
    // rows 0..C-1 = scale*T_c (centre pixel), row C = 1 (softmax denominator), rows C+1.. = residual plane
    const int bt = tid - 64;   // 0..63
    for (int n = 0; n < n_img; ++n) {
      const int s = n % S;
      mbar_wait(bar_full + 8 * s, (n / S) & 1, 6);
      const uint8_t* st = sStage + (size_t)s * g.stage_bytes;
      uint8_t* vt = sStage + (size_t)s * g.stage_bytes + g.vt_off;
      int tt = 0;
      for (int ch = 0; ch < g.nchunks; ++ch) {
        const int N = 8 * g.chunk_g[ch], u0 = g.chunk_u0[ch];
        for (int vb = 0; vb < g.nvb; ++vb, ++tt) {
          uint8_t* vtile = vt + (size_t)tt * g.vt_tile * 4;
          // one thread per (candidate, row group): 16 half values of one candidate = rows 0..7 (group 0) / 8..15 (group 1)
          for (int e = bt; e < 2 * N; e += 64) {
            const int r = e >> 1, grp = e & 1;
            const int u = u0 + (r >> 3), v = 8 * vb + (r & 7);
            const bool valid = (u < g.Ph) && (v < g.Pw);
            __half rows[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) rows[q] = __float2half_rn(0.f);
            if (grp == 0 && valid) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const size_t go = ((size_t)(c * g.R + (u - u0) + g.d) * g.W + (v + g.d)) * 16;
                rows[c] = *reinterpret_cast<const __half*>(st + go);
                if (g.bank_planes > 1) rows[C + 1 + c] = *reinterpret_cast<const __half*>(st + g.img_bytes + g.tile_pad + go);
              }
              rows[C] = __float2half_rn(1.f);
            }
            __half* dst = reinterpret_cast<__half*>(vtile + (size_t)(r >> 3) * 256 + grp * 128) + (r & 7);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q * 8] = rows[q];
          }
        }
      }
      fence_proxy_async();           // generic-proxy writes above are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_vready + 8 * s);
    }
  } else {
    // =========================== epilogue
    // Tile T sits in S buffer T & 1.  Warpgroup wg owns the contiguous chunks [wg*cw, wg*cw + cw) of 16 columns.
    const int wg = (warp - 4) >> 2, q = tid - 128 - wg * 128;   // q = query row = TMEM lane
    const int qi = i0 + (q >> 3), qj = j0 + (q & 7);
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    const float c1 = CDS_LOG2E * a / beta * inv_scale;
    const float2 c1c1 = make_float2(c1, c1);
    float m_run = -INFINITY;                    // common to the four threads that share a query row
    // folded state (warpgroup 0 only)
    float m_acc = -INFINITY, l = 0.f, acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float mref0 = 0.f, mref1 = 0.f;             // reference max used for the P of the tile in each S buffer
    float* dbg = (p.dbg && split == 0 && qi < g.H && qj < g.W)
                     ? p.dbg + ((size_t)b * g.H * g.W + (size_t)qi * g.W + qj) * ((size_t)g.Ph * g.Pw)
                     : nullptr;

    auto fold = [&](int T) {                    // O tile of tile T -> (m_acc, l, acc); warpgroup 0 only
      const int buf = T & 1;
      mbar_wait(bar_oready + 8 * buf, (uint32_t)((T >> 1) & 1), 9);
      tc_fence_after();
      uint32_t o[8];
      tmem_ld8(tmem_base + O_COL0 + buf * 16 + lane_addr, o);
      tmem_ld_wait8(o);
      const float mr = buf ? mref1 : mref0;
      const float mn = fmaxf(m_acc, mr);
      const float s0 = (m_acc == -INFINITY) ? 0.f : ex2(m_acc - mn);
      const float s1 = (mr == -INFINITY) ? 0.f : ex2(mr - mn);
      l = l * s0 + __uint_as_float(o[C]) * s1;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v = __uint_as_float(o[c]);
        if (g.bank_planes > 1) v += __uint_as_float(o[C + 1 + c]);
        acc[c] = acc[c] * s0 + v * s1;
      }
      m_acc = mn;
    };

    int T = 0;
    for (int n = 0; n < n_img; ++n) {
      const float lw = __ldg(p.logw + n0 + n) * CDS_LOG2E;
      const bool dump = (dbg != nullptr) && n == 0;
      for (int ch = 0; ch < g.nchunks; ++ch) {
        const int N = 8 * g.chunk_g[ch], u0 = g.chunk_u0[ch];
        const int nck = N >> 4, cw = (nck + NUM_EPI_WG - 1) / NUM_EPI_WG;
        const int c_lo = 16 * wg * cw, c_hi = min(N, c_lo + 16 * cw);     // this warpgroup's columns
        for (int vb = 0; vb < g.nvb; ++vb, ++T) {
          const int buf = T & 1;
          if (wg == 0 && T >= 2) fold(T - 2);     // frees the O tile of this buffer before P.V(T) can be issued
          mbar_wait(bar_tfull + 8 * buf, (uint32_t)((T >> 1) & 1), 5);
          tc_fence_after();
          const uint32_t taddr = tmem_base + buf * S_BUF_COLS + lane_addr;
          const bool edge = 8 * vb + 8 > g.W;
          const int nval_v = g.Pw - 8 * vb, nval_u = g.Ph - u0;
          // ---- sweep 1: this warpgroup's best logit per query
          float dmax = -INFINITY;
          for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait16(r);
            if (edge) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (((c0 + e) & 7) >= nval_v || ((c0 + e) >> 3) >= nval_u) r[e] = 0xff7fffffu;
            }
#pragma unroll
            for (int e = 0; e < 16; e += 2) dmax = max3(dmax, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
            if (dump) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int u = u0 + ((c0 + e) >> 3), v = 8 * vb + ((c0 + e) & 7);
                if (u < g.Ph && v < g.Pw) dbg[u * g.Pw + v] = __uint_as_float(r[e]) * inv_scale;
              }
            }
          }
          // ---- one reference max per query for the whole tile (P.V sums over all of its columns)
          float* smx = sMax + (T & 1) * (NUM_EPI_WG * 128);
          smx[wg * 128 + q] = (c_lo < c_hi) ? fmaf(dmax, c1, lw) : -INFINITY;
          bar_sync_named(1, 128 * NUM_EPI_WG);
          float m_tile = smx[q];
#pragma unroll
          for (int w = 1; w < NUM_EPI_WG; ++w) m_tile = fmaxf(m_tile, smx[w * 128 + q]);
          m_run = fmaxf(m_run, m_tile);
          if (buf) mref1 = m_run; else mref0 = m_run;
          // ---- sweep 2: P = 2^(logit - m_run + 14) as fp16, written over the consumed S columns
          const float off = lw - m_run + P_SHIFT;
          const float2 off2 = make_float2(off, off);
          const float skip_d = (m_run - SKIP_LOG2 - lw) / c1;     // accumulator value below which a weight is < 2^-40
          int i = 0;
          for (int c0 = c_lo; c0 < c_hi; c0 += 16, ++i) {
            uint32_t r[16], pk[8];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait16(r);
            if (edge) {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (((c0 + e) & 7) >= nval_v || ((c0 + e) >> 3) >= nval_u) r[e] = 0xff7fffffu;
            }
            float cm = max3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
#pragma unroll
            for (int e = 3; e < 15; e += 2) cm = max3(cm, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
            cm = fmaxf(cm, __uint_as_float(r[15]));
            if (__all_sync(0xffffffffu, cm < skip_d)) {
#pragma unroll
              for (int e = 0; e < 8; ++e) pk[e] = 0u;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float2 ar = fma2(make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), c1c1, off2);
                pk[e] = pack_f16x2(ex2(ar.x), ex2(ar.y));
              }
            }
            tmem_st8(taddr + c_lo + 8 * i, pk);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pready + 8 * buf);
        }
      }
    }
    if (wg == 0) {
      if (total_tiles >= 2) fold(total_tiles - 2);
      fold(total_tiles - 1);
      if (qi < g.H && qj < g.W) {
        // undo the 2^14 shift of P and the bank scale of V'
        const float k_l = exp2f(-P_SHIFT), k_a = k_l * inv_scale;
        const int HW = g.H * g.W, pix = qi * g.W + qj;
        const size_t o = ((size_t)split * p.B + b) * HW + pix;
        p.m[o] = m_acc;
        p.l[o] = l * k_l;
#pragma unroll
        for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b) * C + c) * HW + pix] = acc[c] * k_a;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

extern "C" int64_t cds_els_umma_pv_smem_bytes(int C, int H, int W, int k, int passes, int bank_planes) {
  UmmaGeom g;
  uint2 local[MAX_MMAS];
  return make_geom(C, H, W, k, passes, bank_planes, g, local, 1) ? g.smem_total : 0;
}

extern "C" int cds_els_partials_umma_pv(int query_pad, const float* x, int B, int C, int H, int W, int k,
                                        const float* beta, const void* bank_hi, const void* bank_lo, float bank_scale,
                                        const void* norm_plane, const int32_t* idx, const float* logw, int64_t n_sel,
                                        int splits, int passes, float* m, float* l, float* acc, float* dbg_dots,
                                        void* stream) {
  UmmaParams p;
  const int planes = bank_lo ? 2 : 1;
  if (!make_geom(C, H, W, k, passes, planes, p.g, p.table, 1)) {
    cds_set_error("cds_els_partials_umma_pv: unsupported geometry C=%d H=%d W=%d k=%d passes=%d planes=%d", C, H, W, k,
                  passes, planes);
    return CDS_ERR_UNSUPPORTED;
  }
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1, "cds_els_partials_umma_pv: empty problem");
  if (splits > n_sel) splits = (int)n_sel;
  p.B = B; p.pad = query_pad; p.splits = splits; p.n_sel = n_sel;
  p.x = x; p.beta = beta;
  p.bank_hi = (const uint8_t*)bank_hi; p.bank_lo = (const uint8_t*)bank_lo;
  p.norm_plane = (const uint8_t*)norm_plane;
  p.scale = bank_scale;
  p.idx = idx; p.logw = logw;
  p.m = m; p.l = l; p.acc = acc; p.dbg = dbg_dots;
  p.flags = 0;
  const int tiles = ((H + TI - 1) / TI) * ((W + TJ - 1) / TJ);
  dim3 grid(tiles, splits, B);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define LAUNCH(CC)                                                                                                    \
  case CC:                                                                                                            \
    e = cudaFuncSetAttribute(els_umma_pv_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g.smem_total);    \
    if (e == cudaSuccess) els_umma_pv_kernel<CC><<<grid, THREADS, p.g.smem_total, st>>>(p);                           \
    break;
  switch (C) {
    LAUNCH(1) LAUNCH(2) LAUNCH(3)
  }
#undef LAUNCH
  if (e != cudaSuccess) {
    cds_set_error("els_umma_pv_kernel attribute: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  CDS_CHECK_LAUNCH("els_umma_pv_kernel");
  return CDS_OK;
}
