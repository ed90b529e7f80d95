// ELS score partials, "P.V" variant: the softmax-weighted sum of the centre pixels also runs on the tensor cores.
//
// Same contraction, staging and descriptor aliasing as els_umma.cu (see there).  What differs is the epilogue:
//   S tile (128 queries x N <= 192 candidates, fp32 in TMEM)
//     --epilogue-->  P = 2^(logit - m_wg), written back IN PLACE over S (tcgen05.st) and read by the tensor core as tf32
//     --second UMMA (kind::tf32, A from TMEM)-->  O_wg[128 x 16] = P[:, columns of warpgroup wg] . V'[those columns x 16]
//   V' = (scale*T_c ..., 1, residual planes ...) staged per tile as a K-major fp32 operand by the builder warps; column C
//   of O is the softmax denominator.  tf32 keeps fp32's exponent range, so P needs no shift and no clamping; its 11-bit
//   significand rounds every weight by <= 2^-11 relative (numerator and denominator alike).
// Every epilogue warpgroup owns the 16-column chunks wg, wg+4, ... of a tile, its own running max m_wg and its own
// 16-column O tile per S buffer, so nothing is exchanged between warpgroups per tile: sweep 1 = max over the own
// columns, sweep 2 = exp2 + store.  The (m, l, acc) state of a warpgroup is the fold of its O tiles; the four states are
// merged once at the end.  The weighted sum therefore costs no FMA-pipe instruction and no shared-memory load.
// TMEM: S buffers at columns 0 and 192, O tiles at 384 + 16*(4*buffer + warpgroup).
#include "umma_common.cuh"

using namespace umma;

namespace {

constexpr int S_BUF_COLS = 192;
constexpr int O_COL0 = 384;

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
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

  uint8_t* sA = smem;
  uint8_t* sStage = smem + g.smem_A;
  float* sMax = reinterpret_cast<float*>(smem + g.smem_A + g.smem_stage);   // merge scratch of the warpgroups' final states
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

  const int units = n_img * g.nchunks;            // staging units (image, row band)
  const int total_tiles = units * g.nvb;

  if (warp == 0) {
    // =========================== producer: per (image, band) unit one bulk copy per channel (+ residual plane) of the
    // band's strip rows, and one of its norm-plane rows
    if (lane == 0) {
      int unit = 0;
      for (int n = 0; n < n_img; ++n) {
        const long long gi = p.idx[n0 + n];
        for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
          const int s = unit % S, u0 = g.chunk_u0[ch];
          const uint32_t rows = (uint32_t)min(g.R, g.H - u0), nrows = (uint32_t)min(g.G, g.H - u0);
          const uint32_t cbytes = rows * g.S1, nbytes = nrows * g.S1;
          mbar_wait(bar_empty + 8 * s, ((unit / S) & 1) ^ 1, 1);
          const uint32_t dst = smem_u32(sStage + (size_t)s * g.stage_bytes);
          mbar_expect_tx(bar_full + 8 * s, g.bank_planes * g.C * cbytes + nbytes);
          for (int c = 0; c < g.C; ++c) {
            const size_t src = ((size_t)gi * g.C + c) * g.chan_bytes + (size_t)u0 * g.S1;
            bulk_g2s(dst + c * g.R * g.S1, p.bank_hi + src, cbytes, bar_full + 8 * s);
            if (g.bank_planes > 1)
              bulk_g2s(dst + g.img_bytes + g.tile_pad + c * g.R * g.S1, p.bank_lo + src, cbytes, bar_full + 8 * s);
          }
          bulk_g2s(dst + g.np_off, p.norm_plane + (size_t)gi * g.chan_bytes + (size_t)u0 * g.S1, nbytes, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer.  Issue order: S(0), S(1), P.V(0), S(2), P.V(1), S(3), ...  S(T+2) reuses
    // the buffer of tile T and is issued after P.V(T), which itself waits until every epilogue warp has finished
    // reading S(T) and writing P(T) over it; tensor-core operations execute in issue order.
    const uint64_t a_hi = desc_hi(g.RA), b_hi = desc_hi(g.S1);
    const uint64_t v_hi = desc_hi(128);                 // V' operand: 8-row groups 128 B apart
    const uint32_t a_base = smem_u32(sA) >> 4;
    const int nm = g.n_mma;
    const int N = 8 * g.G, nck = N >> 4;                // every band has G patch rows: N columns, nck chunks of 16
    // P.V instruction: D = f32 [4,6)=1, A = B = tf32 [7,10)=[10,13)=2, K-major, N = 16, M = 128; K = 8 per instruction
    const uint32_t idesc_pv = (1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);

    auto issue_pv = [&](int T) {
      const int unit = T / g.nvb, vb = T - unit * g.nvb;
      const int s = unit % S, buf = T & 1;
      if (vb == 0) mbar_wait(bar_vready + 8 * s, (unit / S) & 1, 7);        // V' operands of this unit are built
      mbar_wait(bar_pready + 8 * buf, (uint32_t)((T >> 1) & 1), 8);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vt = (smem_u32(sStage + (size_t)s * g.stage_bytes + g.vt_off) + (uint32_t)vb * g.vt_tile * 4u) >> 4;
        const uint32_t s_col = tmem_base + buf * S_BUF_COLS;
        const uint32_t o_col = tmem_base + O_COL0 + buf * (16 * NUM_EPI_WG);
        for (int j = 0; j < nck; ++j) {
          const int w = j % NUM_EPI_WG;                                       // owner of chunk j
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // 8 candidates = two K granules of 4 fp32 (256 B apart); V' advances 512 B per 8 candidates
            const uint64_t bd = v_hi | (uint64_t)(((vt + (uint32_t)(2 * j + h) * 32u) & 0x3FFFu) | (16u << 16));
            umma_tf32_ts(o_col + 16 * w, s_col + 16 * j + 8 * h, bd, idesc_pv, (j >= NUM_EPI_WG || h) ? 1u : 0u);
          }
        }
        umma_commit(bar_oready + 8 * buf);
        if (vb == g.nvb - 1) umma_commit(bar_empty + 8 * s);                  // last reader of this stage
      }
      __syncwarp();
    };

    int T = 0;
    for (int unit = 0; unit < units; ++unit) {
      const int s = unit % S;
      const uint32_t stage_addr = smem_u32(sStage + (size_t)s * g.stage_bytes);
      for (int vb = 0; vb < g.nvb; ++vb, ++T) {
        // the pending P.V first: it may be the commit that releases the stage the next unit is waiting for
        if (T >= 2) issue_pv(T - 2);
        if (vb == 0) {
          mbar_wait(bar_full + 8 * s, (unit / S) & 1, 2);
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
    if (total_tiles >= 2) issue_pv(total_tiles - 2);
    issue_pv(total_tiles - 1);
  } else if (warp == 2 || warp == 3) {
    // =========================== builders: V' operand of every tile of the staged band, fp32 (read as tf32), K-major
    // no-swizzle: element (row nrow, candidate column r) at (r/4)*256 + (nrow/8)*128 + (nrow%8)*16 + (r%4)*4 bytes;
    // rows 0..C-1 = scale*T_c (centre pixel), row C = 1 (softmax denominator), rows C+1.. = residual plane, rest 0
    const int bt = tid - 64;   // 0..63
    const int N = 8 * g.G;
    int unit = 0;
    for (int n = 0; n < n_img; ++n) {
      for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
        const int s = unit % S, u0 = g.chunk_u0[ch];
        mbar_wait(bar_full + 8 * s, (unit / S) & 1, 6);
        const uint8_t* st = sStage + (size_t)s * g.stage_bytes;
        uint8_t* vt = sStage + (size_t)s * g.stage_bytes + g.vt_off;
        for (int vb = 0; vb < g.nvb; ++vb) {
          uint8_t* vtile = vt + (size_t)vb * g.vt_tile * 4;
          for (int r = bt; r < N; r += 64) {
            const int u = u0 + (r >> 3), v = 8 * vb + (r & 7);
            float rows[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (u < g.Ph && v < g.Pw) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const size_t go = ((size_t)(c * g.R + (u - u0) + g.d) * g.W + (v + g.d)) * 16;
                rows[c] = __half2float(*reinterpret_cast<const __half*>(st + go));
                if (g.bank_planes > 1)
                  rows[C + 1 + c] = __half2float(*reinterpret_cast<const __half*>(st + g.img_bytes + g.tile_pad + go));
              }
              rows[C] = 1.f;
            }
            float* dst = reinterpret_cast<float*>(vtile + (size_t)(r >> 2) * 256) + (r & 3);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              dst[q * 4] = rows[q];            // row group 0 (rows 0..7)
              dst[32 + q * 4] = 0.f;           // row group 1 (rows 8..15) is unused
            }
          }
        }
        fence_proxy_async();           // generic-proxy writes above are read by the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_vready + 8 * s);
      }
    }
  } else {
    // =========================== epilogue: tile T sits in S buffer T & 1; warpgroup wg owns chunks wg, wg+4, ...
    const int wg = (warp - 4) >> 2, q = tid - 128 - wg * 128;   // q = query row = TMEM lane
    const int qi = i0 + (q >> 3), qj = j0 + (q & 7);
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;
    const float c1 = CDS_LOG2E * a / beta * inv_scale;
    const float2 c1c1 = make_float2(c1, c1);
    // the norm-plane marker suppresses invalid positions by 2^-(log2e * a^2/(2 beta) * 3 * INVALID_NORM); when beta -> 1 that
    // factor fades (a -> 0), so fall back to explicit column masking
    const bool weak_marker = CDS_LOG2E * a * a / (2.f * beta) * 3.f * INVALID_NORM < 64.f;
    const int N = 8 * g.G;
    const bool has_cols = 16 * wg < N;                          // a narrow tile may leave the last warpgroups idle
    float m_run = -INFINITY;                                    // running max over this warpgroup's columns
    float m_acc = -INFINITY, l = 0.f, acc[C];                   // fold of this warpgroup's O tiles
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    float mref0 = 0.f, mref1 = 0.f;                             // m_run used for the P of the tile in each S buffer
    float* dbg = (p.dbg && split == 0 && qi < g.H && qj < g.W)
                     ? p.dbg + ((size_t)b * g.H * g.W + (size_t)qi * g.W + qj) * ((size_t)g.Ph * g.Pw)
                     : nullptr;

    auto fold = [&](int T) {                                    // O tile of (tile T, this warpgroup) -> (m_acc, l, acc)
      const int buf = T & 1;
      mbar_wait(bar_oready + 8 * buf, (uint32_t)((T >> 1) & 1), 9);
      tc_fence_after();
      if (!has_cols) return;
      uint32_t o[8];
      tmem_ld8(tmem_base + O_COL0 + buf * (16 * NUM_EPI_WG) + 16 * wg + lane_addr, o);
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

    int T = 0, unit = 0;
    for (int n = 0; n < n_img; ++n) {
      const float lw = __ldg(p.logw + n0 + n) * CDS_LOG2E;
      const bool dump = (dbg != nullptr) && n == 0;
      for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
        const int u0 = g.chunk_u0[ch];
        for (int vb = 0; vb < g.nvb; ++vb, ++T) {
          const int buf = T & 1;
          if (T >= 2) fold(T - 2);                // frees this warpgroup's O tile of the buffer before P.V(T) is issued
          mbar_wait(bar_tfull + 8 * buf, (uint32_t)((T >> 1) & 1), 5);
          tc_fence_after();
          const uint32_t taddr = tmem_base + buf * S_BUF_COLS + lane_addr;
          const bool edge = weak_marker || 8 * vb + 8 > g.W;
          const int nval_v = g.Pw - 8 * vb, nval_u = g.Ph - u0;
          // ---- sweep 1: best logit of this warpgroup's columns (c1 > 0, so the max commutes with the affine map)
          float dmax = -INFINITY;
          for (int c0 = 16 * wg; c0 < N; c0 += 16 * NUM_EPI_WG) {
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
          if (has_cols) m_run = fmaxf(m_run, fmaf(dmax, c1, lw));
          if (buf) mref1 = m_run; else mref0 = m_run;
          // ---- sweep 2: P = 2^(logit - m_run), stored over the S columns it came from
          const float off = (m_run == -INFINITY) ? -INFINITY : lw - m_run;   // nothing valid seen yet: all weights 0
          const float2 off2 = make_float2(off, off);
          const float skip_d = (m_run - SKIP_LOG2 - lw) / c1;     // accumulator value below which a weight is < 2^-40
          for (int c0 = 16 * wg; c0 < N; c0 += 16 * NUM_EPI_WG) {
            uint32_t r[16];
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
              for (int e = 0; e < 16; ++e) r[e] = 0u;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float2 ar = fma2(make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), c1c1, off2);
                r[2 * e] = __float_as_uint(ex2(ar.x));
                r[2 * e + 1] = __float_as_uint(ex2(ar.y));
              }
            }
            tmem_st16(taddr + c0, r);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pready + 8 * buf);
        }
      }
    }
    if (total_tiles >= 2) fold(total_tiles - 2);
    fold(total_tiles - 1);
    // undo the bank scale of V', merge the warpgroups' states and write this split's partials
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] *= inv_scale;
    float* sMerge = sMax;                                    // [NUM_EPI_WG-1][128][2+C]
    if (wg > 0) {
      float* dst = sMerge + ((wg - 1) * 128 + q) * (2 + C);
      dst[0] = m_acc;
      dst[1] = l;
#pragma unroll
      for (int c = 0; c < C; ++c) dst[2 + c] = acc[c];
    }
    bar_sync_named(1, 128 * NUM_EPI_WG);
    if (wg == 0 && qi < g.H && qj < g.W) {
      float M = m_acc;
#pragma unroll
      for (int w = 1; w < NUM_EPI_WG; ++w) M = fmaxf(M, sMerge[((w - 1) * 128 + q) * (2 + C)]);
      const float w0 = (m_acc == -INFINITY) ? 0.f : ex2(m_acc - M);
      float L = l * w0, A[C];
#pragma unroll
      for (int c = 0; c < C; ++c) A[c] = acc[c] * w0;
#pragma unroll
      for (int w = 1; w < NUM_EPI_WG; ++w) {
        const float* src = sMerge + ((w - 1) * 128 + q) * (2 + C);
        const float ww = (src[0] == -INFINITY) ? 0.f : ex2(src[0] - M);
        L = fmaf(src[1], ww, L);
#pragma unroll
        for (int c = 0; c < C; ++c) A[c] = fmaf(src[2 + c], ww, A[c]);
      }
      const int HW = g.H * g.W, pix = qi * g.W + qj;
      const size_t o = ((size_t)split * p.B + b) * HW + pix;
      p.m[o] = M;
      p.l[o] = L;
#pragma unroll
      for (int c = 0; c < C; ++c) p.acc[(((size_t)split * p.B + b) * C + c) * HW + pix] = A[c];
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
  p.bank_rows = nullptr;
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
