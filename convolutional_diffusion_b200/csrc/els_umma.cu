// ELS score partials on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators) with implicit im2col.
//
// What it computes (reference: /root/reference/src/utils/idealscore.py:397-473, restated):
//   for every query pixel (i,j) of x (k x k x C patch of the padded image) and every valid k x k x C patch
//   p(n,u,v) of every selected bank image:  logit = -(|q|^2 - 2a q.p + a^2 |p|^2)/(2 beta) + logw_n,
//   online softmax over all (n,u,v), weighted sum of the patch-centre pixels.  |q|^2 is constant per query
//   and cancels in the softmax, so it is dropped from the logits (the partial max m is in those units).
//
// How (B200-first, not a translation of unfold+conv2d):
//   * The bank lives in HBM as "strip8": one 16-byte granule = 8 vertically adjacent pixels
//     [n][c][u][x][8] fp16.  With K-major, no-swizzle UMMA descriptors the row pitch inside an 8-row core
//     matrix is 16 bytes, so 8 consecutive B rows = 8 horizontally adjacent patches, the next K granule
//     (LBO) = the next patch column dx, and the 8-row-group stride (SBO = one image row of granules) = the next
//     patch row u.  Overlapping patches therefore alias the SAME shared-memory bytes: patches are never
//     materialised, in HBM or in shared memory.  A whole image is staged by one cp.async.bulk.
//   * The query side uses the same trick on a per-CTA tile of 16 x 8 pixels, with the K padding (dy >= k)
//     zeroed on the query side only, so the bank can stay unmasked and k-independent.
//   * The -a^2|p|^2/(2 beta) term rides in the SAME contraction: one extra K granule per patch holds |p|^2 as a
//     three-way fp16 split (the per-k "norm plane", same strip addressing), its query-side partner holds the
//     constant -a*scale/2, so the accumulator is already  scale*(q.p) - (a*scale/2)|p|^2  and the logit is a
//     single multiply away.  Invalid patch positions carry a huge norm and vanish in the softmax.
//   * One UMMA covers N = 8 * (#patch rows) <= 256 candidates x 128 queries x K = 16; accumulators ping-pong
//     between two TMEM buffers.  Warp roles: producer (bulk copies), MMA issuer (one thread), two builder warps
//     (centre-pixel tables per image) and four epilogue warpgroups, decoupled through mbarriers.
//   * Mixed K layout (k > 8, k % 8 != 0, single-plane bank): the trailing k % 8 patch rows are contracted as horizontal
//     8-pixel granules from the "rows8" plane instead of one more mostly-empty vertical block (k=9: 17 UMMAs per tile
//     instead of 28).  Their query slices cannot alias, so they live only in TMEM (built once in the staging area).
//   * Query K slices stay resident in the TMEM columns the two accumulator buffers leave free: those UMMAs read A from
//     TMEM and only B from shared memory (the data pipe they share with the epilogue's loads was 95 % busy).
//   * Epilogue = flash softmax in two passes per 16-column chunk on the 16x256b TMEM fragment (four query rows x four
//     columns per thread, so one 16-byte centre-pixel load per channel serves 16 pairs): (1) tcgen05.ld + 3-input max;
//     a chunk whose best logit is 2^-40 below the running max for every row of the warp is skipped; (2) packed f32x2
//     fma -> ex2 -> weighted sums.  Floor = 8 cycles per candidate column twice over (TMEM read bandwidth, MUFU).
//   * P.V epilogue (template flag PV, single-plane banks): the weighted sum of the centre pixels is a second contraction on
//     the tensor cores.  One sweep over the accumulators: tcgen05.ld (row per thread), row max, P = 2^(logit - m_ref + 8)
//     packed to fp16 and stored IN PLACE over the consumed columns (tcgen05.st); the MMA warp then issues
//     O[128 x 16] += P[128 x 16 candidates] . V'[16 candidates x 16] per 16-column chunk with P as the TMEM-resident A
//     operand.  V' (centre pixels and a ones row, fp16, exact for 8-bit banks) is routed per chunk to the 4 O columns of
//     the epilogue warpgroup that owns the chunk, so every (warpgroup, row) has a private reference m_ref and a private
//     accumulator and nothing is exchanged per tile.  m_ref is not a running max: it is re-based on a fixed schedule
//     (before tile 1, 2, 4, 8, ... of the CTA's slice: drain O into the thread's fp32 state, m_ref = best logit seen);
//     a chunk holding a logit more than 2^7.5 above m_ref (fp16 overflow) is evaluated exactly on the CUDA cores into the
//     same fp32 state and stores P = 0, so the result never depends on how good the reference is -- only the speed does.
//     Measured basis (profiles/r02_tmem_ld_floor.log): tcgen05.ld 32x32b delivers 700-800 B/clk/SM (the 16x256b shape
//     250), MUFU.EX2 16/clk/SM, an N=16 UMMA ~29 clk.
#include <type_traits>
#include "umma_common.cuh"

// P.V epilogue: which column pairs of a 32-column unit form their weights with the FMA-pipe polynomial instead of MUFU.EX2.
// Measured (profiles/r02q_*): per tile the epilogue costs a fixed ~750 cycles plus max(issue slots, MUFU cycles); with every
// exponential on MUFU (16 lanes/clk/SM = 8 cycles per column) both are ~1900 cycles for 240 columns, so the polynomial only
// pays together with the wide units that cut the instruction count.
#ifndef CDS_PV_ABLATE
#define CDS_PV_ABLATE 0            // A/B builds only (wrong results): see the uses
#endif
#ifndef CDS_PV_POLY_MASK
#define CDS_PV_POLY_MASK 0x4444    // pairs 2, 6, 10, 14: a quarter of the exponentials
#endif

using namespace umma;

namespace {

#ifdef CDS_PROFILE_SWITCHES
// profile builds only: [0] warp-chunks seen, [1] warp-chunks skipped (all weights < 2^-40 of the running max),
// [2] warp-chunks that raised a running max, [3] (warp, tile) units seen, [4] (warp, tile) units with every chunk skipped
// [8..10] MMA warp of the P.V variant, clock cycles summed over CTAs: waiting for barriers, issuing, tiles; [11..13] the same
// for epilogue warp 0: waiting for the accumulators, working, tiles
__device__ unsigned long long g_els_counters[16];
#endif

// ------------------------------------------------------------------ the kernel
template <int C, bool PV>
__global__ void __launch_bounds__(THREADS, 1) els_umma_kernel(const __grid_constant__ UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const UmmaGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Warp roles.  The warp scheduler prefers the highest warp id among the eligible warps of a sub-partition, and the MMA warp
  // must never wait for an issue slot (the tensor pipe queues only 2-3 UMMAs), so it takes the last warp; the epilogue
  // warps come first (TMEM lane quadrant = warp % 4 either way).
#ifndef CDS_ROLES_HEAD
  constexpr int W_EPI0 = 0, W_PROD = 4 * NUM_EPI_WG, W_B0 = 4 * NUM_EPI_WG + 1, W_MMA = 4 * NUM_EPI_WG + 3;
#else
  constexpr int W_PROD = 0, W_MMA = 1, W_B0 = 2, W_EPI0 = 4;     // A/B: the round-1 assignment
#endif
  // query window: the tiles cover the pixels [qi0, qi0+qrows) x [qj0, qj0+qcols) (the whole image, or the bbELS centre)
  const int tiles_j = (p.qcols + TJ - 1) / TJ;
  const int i0 = p.qi0 + (blockIdx.x / tiles_j) * TI, j0 = p.qj0 + (blockIdx.x % tiles_j) * TJ;
  const int split = blockIdx.y, b = blockIdx.z;
  const long long n0 = p.n_sel * split / p.splits, n1 = p.n_sel * (split + 1) / p.splits;
  const int n_img = (int)(n1 - n0);
  const int S = g.stages;   // 1 or 2: stage = unit & (S-1), phase = (unit >> (S-1)) & 1

  uint8_t* sA = smem;
  uint8_t* sStage = smem + g.smem_A;
  float* sMerge = reinterpret_cast<float*>(sA);           // [NUM_EPI_WG-1][128][2+C], aliases the query tile (see make_geom)
  uint64_t* sBar = reinterpret_cast<uint64_t*>(smem + g.smem_A + g.smem_stage + g.smem_merge + g.smem_table);
  // barriers: full[2], empty[2], vready[2], tfull[2], tempty[2] (P.V: pready[2]), pvdone[2]; then the TMEM base address
  const uint32_t bar_full = smem_u32(sBar), bar_empty = bar_full + 16, bar_vready = bar_full + 32;
  const uint32_t bar_tfull = bar_full + 48, bar_tempty = bar_full + 64, bar_pvdone = bar_full + 80;
  const uint32_t bar_pready = bar_tempty, bar_vempty = bar_full + 96;
  uint32_t* sTmemBase = reinterpret_cast<uint32_t*>(sBar + 14);

  const float beta = p.beta[b];
  const float a = sqrtf(1.f - beta);
  const float inv_scale = 1.f / p.scale;

  // ---- one-time setup: barriers, TMEM, descriptor table, zero guards, query tile (A operand)
  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, PV ? 1 + 2 : 1 + 4 * NUM_EPI_WG);   // MMA commit + every epilogue warp (P.V: + the two
                                                                        // builder warps, the only other readers of a stage)
      mbar_init(bar_vempty + 8 * s, 1);                   // P.V: commit after the band's last P.V releases its V' slot
      mbar_init(bar_vready + 8 * s, 2);                   // the two builder warps
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(bar_tfull + 8 * q, 1);
      mbar_init(bar_tempty + 8 * q, 4 * NUM_EPI_WG);      // every epilogue warp reads every tile (P.V: has stored its P)
      mbar_init(bar_pvdone + 8 * q, 1);                   // P.V: [0] = commit after the very last P.V
    }
    fence_barrier_init();
  }
  if (warp == W_B0) tmem_alloc(smem_u32(sTmemBase), TMEM_COLS);
  // zero the staging area once: guard pads behind the tiles and band rows past the image bottom are never written by
  // the bulk copies, and whatever masked columns / zero-weighted K slots read there must stay finite
  for (int e = tid * 16; e < g.smem_stage; e += THREADS * 16) *reinterpret_cast<uint4*>(sStage + e) = make_uint4(0, 0, 0, 0);
  if (g.rem) __syncthreads();     // the horizontal query regions below are built inside the zeroed staging area
  {
    // A[plane][c][blk][gi][jj][8]: element e of granule (gi,jj) of block blk = xpad[i0+gi+8*blk+e][j0+jj] (patch
    // coordinates: padded row = pixel row + dy, i.e. source pixel row i0+gi+dy-d), zero where dy = 8*blk+e >= k
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
    // mixed K layout: patch rows dy = 8*nb + r (r < rem) as horizontal granules; region (c,r,blk)[gi][jj][8], element e of
    // granule (gi,jj) = xpad[i0+gi+dy][j0+jj+8*blk+e], zero where dx = 8*blk+e >= k.  Scratch copy in the staging area:
    // these slices go to TMEM below and are never read from shared memory by a UMMA
    const int per_h = C * g.rem * g.nbh * TI * 8;
    for (int e = tid; e < per_h; e += THREADS) {
      const int jj = e & 7, gi = (e >> 3) % TI, blk = (e / (8 * TI)) % g.nbh, r = (e / (8 * TI * g.nbh)) % g.rem,
                c = e / (8 * TI * g.nbh * g.rem);
      __half hi[8], lo[8];
      int yr = i0 + gi + 8 * g.nb + r - g.d;
      bool rowok = true;
      if (p.pad == CDS_PAD_CIRCULAR) yr = ((yr % g.H) + g.H) % g.H;
      else rowok = (yr >= 0 && yr < g.H);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int dx = 8 * blk + q;
        int xc = j0 + jj + dx - g.d;
        bool ok = rowok && dx < g.k;
        if (p.pad == CDS_PAD_CIRCULAR) xc = ((xc % g.W) + g.W) % g.W;
        else ok = ok && (xc >= 0 && xc < g.W);
        const float v = ok ? xb[(c * g.H + yr) * g.W + xc] : 0.f;
        hi[q] = __float2half_rn(v);
        lo[q] = __float2half_rn(v - __half2float(hi[q]));
      }
      const size_t off = (size_t)((c * g.rem + r) * g.nbh + blk) * (TI * 128) + (size_t)gi * 128 + (size_t)jj * 16;
      *reinterpret_cast<uint4*>(sStage + off) = *reinterpret_cast<uint4*>(hi);
      if (g.passes > 1) *reinterpret_cast<uint4*>(sStage + g.h_plane + off) = *reinterpret_cast<uint4*>(lo);
    }
    // constant block: every granule = three-way fp16 split of gamma = -a*scale/2, laid out against the norm plane's
    // (ph, pm, pl, ph, pm, ph, 0, 0) so that the K sum is (gh+gm+gl)*(ph+pm+pl) up to terms below 2^-30
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

  // ---- resident query slices: the first n_tmem K=16 slices of the A operand are copied into TMEM once (row q = lane q,
  // one column = two K elements), so those UMMAs stop re-reading 4 KB of shared memory per candidate tile
  if (PV) {          // O accumulator starts at zero: every P.V accumulates
    if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
      const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      tmem_st16(tmem_base + g.o_col + (((uint32_t)((warp & 3) * 32)) << 16), z);
      tmem_st_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (g.n_tmem > 0) {
    if (warp >= W_EPI0 && warp < W_EPI0 + 4) {
      const int q = tid - 32 * W_EPI0;
      const uint32_t dst = tmem_base + g.a_tmem_col + (((uint32_t)((warp & 3) * 32)) << 16);
      for (int t = 0; t < g.n_tmem; ++t) {
        const uint32_t ex = p.table[t].x;
        const uint8_t* rowp = ((ex >> 31) ? sStage + (size_t)(q >> 3) * 128 : sA + (size_t)(q >> 3) * g.RA) + (size_t)(q & 7) * 16;
        const uint4 k0 = *reinterpret_cast<const uint4*>(rowp + ((ex & 0x3FFFu) << 4));
        const uint4 k1 = *reinterpret_cast<const uint4*>(rowp + ((ex & 0x3FFFu) << 4) + (((ex >> 16) & 0x3FFFu) << 4));
        const uint32_t v[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
        tmem_st8(dst + 8 * t, v);
      }
      tmem_st_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (g.rem) {      // give the scratch region back to the pipeline as zeros
      for (int e = tid * 16; e < g.passes * g.h_plane; e += THREADS * 16)
        *reinterpret_cast<uint4*>(sStage + e) = make_uint4(0, 0, 0, 0);
      fence_proxy_async();
      __syncthreads();
    }
  }

  if (warp == W_PROD) {
    // =========================== producer: per (image, band) unit one bulk copy per channel (+ residual plane) of the
    // band's strip rows, and one of its norm-plane rows
    if (lane == 0) {
      int unit = 0;
      for (int n = 0; n < n_img; ++n) {
        const long long gi = p.idx[n0 + n];
        for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
          const int s = unit & (S - 1), u0 = g.chunk_u0[ch];
          const uint32_t rows = (uint32_t)min(g.R, g.H - u0), nrows = (uint32_t)min(g.G, g.H - u0);
          const uint32_t cbytes = rows * g.S1, nbytes = nrows * g.S1;
          mbar_wait(bar_empty + 8 * s, ((unit >> (S - 1)) & 1) ^ 1, 1);
          const uint32_t dst = smem_u32(sStage + (size_t)s * g.stage_bytes);
          const uint32_t hrow0 = (uint32_t)(u0 + 8 * g.nb);
          const uint32_t hbytes = g.rem ? (uint32_t)min(g.Rh, g.H - (int)hrow0) * g.S1 : 0u;
          mbar_expect_tx(bar_full + 8 * s, g.bank_planes * g.C * cbytes + g.C * hbytes + nbytes);
          for (int c = 0; c < g.C; ++c) {
            const size_t src = ((size_t)gi * g.C + c) * g.chan_bytes + (size_t)u0 * g.S1;
            bulk_g2s(dst + c * g.R * g.S1, p.bank_hi + src, cbytes, bar_full + 8 * s);
            if (g.bank_planes > 1)
              bulk_g2s(dst + g.img_bytes + g.tile_pad + c * g.R * g.S1, p.bank_lo + src, cbytes, bar_full + 8 * s);
            if (g.rem)
              bulk_g2s(dst + g.hb_off + c * g.Rh * g.S1, p.bank_rows + ((size_t)gi * g.C + c) * g.chan_bytes + (size_t)hrow0 * g.S1,
                       hbytes, bar_full + 8 * s);
          }
          bulk_g2s(dst + g.np_off, p.norm_plane + (size_t)gi * g.chan_bytes + (size_t)u0 * g.S1, nbytes, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == W_MMA) {
    // =========================== MMA issuer: the warp stays converged, one elected lane issues.  The descriptor
    // table is read from the kernel parameters (constant bank) with warp-uniform indices, so descriptor arithmetic
    // runs on the uniform datapath.
    const uint64_t a_hi = desc_hi(g.RA), b_hi = desc_hi(g.S1);
    const uint32_t a_base = smem_u32(sA) >> 4;
    const int nm = g.n_mma, nt = g.n_tmem;
    const uint32_t a_tmem = tmem_base + g.a_tmem_col;
    if constexpr (PV) {
      // Issue order: S(0), S(1), P.V(0), S(2), P.V(1), ...  S(T+2) reuses the buffer of tile T and is issued after P.V(T),
      // which itself waits until every epilogue warp has read S(T) and stored P(T) over it; tensor-core operations
      // execute in issue order, so no barrier is needed between P.V(T) reading the buffer and S(T+2) overwriting it.
      // The tensor pipe queues only two or three UMMAs ahead of the one it executes (measured: the issue of 10 UMMAs
      // takes as long as their execution), so whatever this warp does between tiles is a bubble: one elected block per
      // tile, descriptors from constant-bank tables, no divisions, no commit per P.V.
      const int nvb = g.nvb, nck = g.G >> 1;              // 16-column chunks per tile (every band has N = 8*G columns)
      const uint32_t idesc = (1u << 4) | (((uint32_t)(8 * g.G) >> 3) << 17) | ((128u >> 4) << 24);
      // P.V instruction: D = f32, A = B = f16, K-major, N = 16, M = 128, K = 16
      const uint32_t idesc_pv = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
      const int units = n_img * g.nchunks, total_tiles = units * nvb;
      const uint32_t o_col = tmem_base + g.o_col;
      const uint32_t vring = smem_u32(sStage + g.v_ring_off) >> 4, vslot = (uint32_t)g.v_slot_bytes >> 4, vtile = (uint32_t)g.v_tile_bytes >> 4;
      const uint32_t stage0 = smem_u32(sStage) >> 4, stage_sz = (uint32_t)g.stage_bytes >> 4;
      int un = 0, vb = 0;                                 // band / tile-in-band of S(T)
      int un2 = 0, vb2 = 0;                               // the same for P.V(T-2)
#ifdef CDS_PROFILE_SWITCHES
      long long ck_wait = 0, ck_issue = 0;
#endif
      for (int T = 0; T < total_tiles + 2; ++T) {
        const bool do_pv = T >= 2, do_s = T < total_tiles;
        const int s = un & 1, s2 = un2 & 1;
        const uint32_t buf = (uint32_t)T & 1u;
#ifdef CDS_PROFILE_SWITCHES
        const long long ck0 = clock64();
#endif
        if (do_pv) {
          if (vb2 == 0) mbar_wait(bar_vready + 8 * s2, (un2 >> 1) & 1, 7);          // V' operands of that band are built
          mbar_wait(bar_pready + 8 * buf, (uint32_t)(((T - 2) >> 1) & 1), 8);       // P(T-2) is stored
        }
        if (do_s && vb == 0) mbar_wait(bar_full + 8 * s, (un >> 1) & 1, 2);
#ifdef CDS_PROFILE_SWITCHES
        const long long ck1 = clock64();
        ck_wait += ck1 - ck0;
#endif
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + buf * g.tmem_buf1;                     // S(T) and P(T-2) share the buffer
#ifdef CDS_PROFILE_SWITCHES
          if (do_pv && !(p.flags & 16)) {   // profiling: 16 = no P.V UMMAs
#else
          if (do_pv) {
#endif
            const uint32_t vt = vring + s2 * vslot + vb2 * vtile;
#pragma unroll 4
            for (int j = 0; j < nck; ++j) {
              const uint2 e = p.pv_table[j];
              umma_f16_ts(o_col, d_tmem + 16 * j, ((uint64_t)e.y << 32) | (uint64_t)(e.x + vt), idesc_pv, 1u);
            }
          }
          if (do_pv && vb2 == nvb - 1) umma_commit(bar_vempty + 8 * s2);            // last reader of that V' slot
          if (do_s) {
            const uint32_t b_base = stage0 + s * stage_sz + vb * 8u;
            int t = 0;
#ifdef CDS_PROFILE_SWITCHES
            if (p.flags & 8) t = nm;      // profiling: 8 = no main UMMAs (accumulators keep whatever they held)
#endif
#pragma unroll 4
            for (; t < nt; ++t)
              umma_f16_ts(d_tmem, a_tmem + 8 * t, b_hi | (uint64_t)(p.table[t].y + b_base), idesc, t ? 1u : 0u);
#pragma unroll 4
            for (; t < nm; ++t) {
              const uint2 e = p.table[t];
              umma_f16(d_tmem, a_hi | (uint64_t)(e.x + a_base), b_hi | (uint64_t)(e.y + b_base), idesc, t ? 1u : 0u);
            }
            umma_commit(bar_tfull + 8 * buf);
            if (vb == nvb - 1) umma_commit(bar_empty + 8 * s);     // the stage's last main UMMAs: back to the producer
          } else if (T == total_tiles + 1) {
            umma_commit(bar_pvdone);      // after the very last P.V: the final drain of the epilogue waits for this
          }
        }
        __syncwarp();
#ifdef CDS_PROFILE_SWITCHES
        ck_issue += clock64() - ck1;
#endif
        if (do_s && ++vb == nvb) { vb = 0; ++un; }
        if (do_pv && ++vb2 == nvb) { vb2 = 0; ++un2; }
      }
#ifdef CDS_PROFILE_SWITCHES
      if (lane == 0) {
        atomicAdd(&g_els_counters[8], (unsigned long long)ck_wait);
        atomicAdd(&g_els_counters[9], (unsigned long long)ck_issue);
        atomicAdd(&g_els_counters[10], (unsigned long long)total_tiles);
      }
#endif
      // nothing may still be executing on the tensor pipe when the CTA gives its TMEM back
      mbar_wait(bar_pvdone, 0u, 10);
    } else {
    long long T = 0;
    int unit = 0;
    for (int n = 0; n < n_img; ++n) {
      for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
        const int s = unit & (S - 1);
        mbar_wait(bar_full + 8 * s, (unit >> (S - 1)) & 1, 2);
        tc_fence_after();
        const uint32_t stage_addr = smem_u32(sStage + (size_t)s * g.stage_bytes);
        const uint32_t N = 8u * g.chunk_g[ch];
        // instruction descriptor: D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0, K-major both,
        // N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | ((N >> 3) << 17) | ((128u >> 4) << 24);
        for (int vb = 0; vb < g.nvb; ++vb, ++T) {
          const int buf = (int)(T & 1);
          mbar_wait(bar_tempty + 8 * buf, (uint32_t)(((T >> 1) & 1) ^ 1), 3);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * g.tmem_buf1;
          const uint32_t b_base = (stage_addr + vb * 128u) >> 4;      // band-relative: patch row u0 is strip row 0
#ifdef CDS_PROFILE_SWITCHES
          if (p.flags & 8) {       // profiling: epilogue only (no UMMAs issued, accumulators keep whatever they held)
            if (elect_one()) umma_commit(bar_tfull + 8 * buf);
            __syncwarp();
            continue;
          }
#endif
          if (elect_one()) {
            int t = 0;
#pragma unroll 4
            for (; t < nt; ++t)       // query slice resident in TMEM
              umma_f16_ts(d_tmem, a_tmem + 8 * t, b_hi | (uint64_t)(p.table[t].y + b_base), idesc, t ? 1u : 0u);
#pragma unroll 4
            for (; t < nm; ++t) {     // query slice read from shared memory
              const uint2 e = p.table[t];
              umma_f16(d_tmem, a_hi | (uint64_t)(e.x + a_base), b_hi | (uint64_t)(e.y + b_base), idesc, t ? 1u : 0u);
            }
            umma_commit(bar_tfull + 8 * buf);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(bar_empty + 8 * s);
        __syncwarp();
      }
    }
    }
  } else if (warp == W_B0 || warp == W_B0 + 1) {
    // =========================== builders: centre pixels of every candidate of the staged image, in tile order:
    // tile (ch,vb), column r = 8*gr + rr <-> patch (u0+gr, 8*vb+rr); pairs of columns are stored side by side
    const int bt = tid - 32 * W_B0;   // 0..63
    int unit = 0;
    if constexpr (PV) {
      // V' operand of the P.V contraction: per tile and patch row kg (= 8 candidates) one 128-byte core matrix of 8 rows x 8
      // fp16.  Chunk j = patch rows 2j, 2j+1 belongs to warpgroup w = (j >> 1) & 3; rows 4*(w&1) + c of its blocks = scale * centre
      // pixel of channel c (the bank's own fp16 values), row 4*(w&1) + 3 = 1 for valid candidates (softmax denominator), everything
      // else stays 0 from the initial clear.  One 16-byte load of strip granule (c, band row 8*b8 + d, x) yields the centre
      // pixels of the eight patch rows kg = 8*b8 .. 8*b8+7 at image column x, i.e. of candidate rr = (x-d) & 7 of tile (x-d) >> 3.
      const int nb8 = (g.G + 7) >> 3;
      for (int n = 0; n < n_img; ++n) {
        for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
          const int s = unit & 1, u0 = g.chunk_u0[ch];
          mbar_wait(bar_vempty + 8 * s, ((unit >> 1) & 1) ^ 1, 11);      // the band two before this one is done with the slot
          mbar_wait(bar_full + 8 * s, (unit >> 1) & 1, 6);
          const uint8_t* st = sStage + (size_t)s * g.stage_bytes;
          uint8_t* vd = sStage + g.v_ring_off + (size_t)s * g.v_slot_bytes + 256;    // operand blocks of tile 0
          const int nrow = min(g.G, g.Ph - u0);                           // valid patch rows of this band
#ifdef CDS_PROFILE_SWITCHES
          if (!(p.flags & 256))       // profiling: 256 = the builders build nothing
#endif
          for (int item = bt; item < C * nb8 * g.Pw; item += 64) {
            const int v = item % g.Pw, cb = item / g.Pw, b8 = cb % nb8, c = cb / nb8;
            const uint4 gr = *reinterpret_cast<const uint4*>(st + ((size_t)(c * g.R + 8 * b8 + g.d) * g.W + (v + g.d)) * 16);
            const uint32_t w4[4] = {gr.x, gr.y, gr.z, gr.w};
            uint8_t* dst = vd + (size_t)(v >> 3) * g.v_tile_bytes + (size_t)(v & 7) * 2 + (size_t)c * 16;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int kg = 8 * b8 + e;
              if (kg < nrow)
                *reinterpret_cast<uint16_t*>(dst + (size_t)kg * 128 + (size_t)(4 * ((kg >> 2) & 1)) * 16) =
                    (uint16_t)(w4[e >> 1] >> (16 * (e & 1)));
            }
          }
          for (int item = bt; item < g.nvb * g.G; item += 64) {           // ones rows (validity can differ between bands)
            const int vb = item / g.G, kg = item - vb * g.G;
            uint32_t o4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              o4[q] = (kg < nrow && 8 * vb + 2 * q < g.Pw ? 0x3C00u : 0u) | (kg < nrow && 8 * vb + 2 * q + 1 < g.Pw ? 0x3C000000u : 0u);
            *reinterpret_cast<uint4*>(vd + (size_t)vb * g.v_tile_bytes + (size_t)kg * 128 + (size_t)(4 * ((kg >> 2) & 1) + 3) * 16) =
                make_uint4(o4[0], o4[1], o4[2], o4[3]);
          }
          if (nrow < g.G) {     // a partial last band: rows that were valid in the previous band of this slot must read as zero
            for (int item = bt; item < g.nvb * (g.G - nrow) * 4; item += 64) {
              const int c = item & 3, rest = item >> 2, vb = rest / (g.G - nrow), kg = nrow + rest % (g.G - nrow);
              *reinterpret_cast<uint4*>(vd + (size_t)vb * g.v_tile_bytes + (size_t)kg * 128 + (size_t)(4 * ((kg >> 2) & 1) + c) * 16) =
                  make_uint4(0, 0, 0, 0);
            }
          }
          fence_proxy_async();           // generic-proxy writes above are read by the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_vready + 8 * s);
            mbar_arrive(bar_empty + 8 * s);   // done reading the stage
          }
        }
      }
    } else
    for (int n = 0; n < n_img; ++n) {
      for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
        const int s = unit & (S - 1);
        mbar_wait(bar_full + 8 * s, (unit >> (S - 1)) & 1, 6);
        const uint8_t* st = sStage + (size_t)s * g.stage_bytes;
        float* vt = reinterpret_cast<float*>(sStage + (size_t)s * g.stage_bytes + g.vt_off);
        const int N = 8 * g.chunk_g[ch], u0 = g.chunk_u0[ch];
        for (int vb = 0; vb < g.nvb; ++vb) {
          float* vtab = vt + (size_t)vb * g.vt_tile;   // [16-column chunk][channel][t4][column pair h][lo]
          for (int r = bt; r < N; r += 64) {
            const int u = u0 + (r >> 3), v = 8 * vb + (r & 7);
            float vals[3] = {0.f, 0.f, 0.f};
            if (u < g.Ph && v < g.Pw) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const size_t go = ((size_t)(c * g.R + (u - u0) + g.d) * g.W + (v + g.d)) * 16;   // band-relative row
                float t = __half2float(*reinterpret_cast<const __half*>(st + go));
                if (g.bank_planes > 1)
                  t += __half2float(*reinterpret_cast<const __half*>(st + g.img_bytes + g.tile_pad + go));
                vals[c] = t * inv_scale;
              }
            }
            const int slot = ((r & 7) >> 1) * 4 + ((r >> 3) & 1) * 2 + (r & 1);
#pragma unroll
            for (int c = 0; c < C; ++c) vtab[((r >> 4) * C + c) * 16 + slot] = vals[c];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_vready + 8 * s);
      }
    }
  } else {
    // =========================== epilogue
    // Tile T sits in TMEM buffer T & 1; while the four warpgroups drain it (warpgroup wg takes the 16-column chunks
    // wg, wg+4, ...), the MMA warp fills the other buffer with tile T+1.
    // TMEM is read with the 16x256b shape: per 16-column chunk a thread holds FOUR query rows (tr, tr+8, tr+16, tr+24 of
    // its warp's 32 lanes) x FOUR columns (the pairs 2*t4 and 8+2*t4), so one 16-byte load of centre pixels per channel
    // serves 16 (query, candidate) pairs: a quarter of the shared-memory traffic of the row-per-thread 32x32b shape, and
    // shared-memory bandwidth (these loads + the UMMA operand reads) is what bounds the kernel.
    const int wg = (warp - W_EPI0) >> 2, lq = warp & 3;
    const uint32_t lane_addr = ((uint32_t)(lq * 32)) << 16;
    const float c1 = CDS_LOG2E * a / beta * inv_scale;   // accumulator -> log2-unit logit
    const float2 c1c1 = make_float2(c1, c1);
    // what both epilogues hand to the merge below: the softmax state of query row q over this warpgroup's columns
    float m = -INFINITY, l = 0.f, acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    int q = 0;
    if constexpr (PV) {
      // ===== P.V epilogue: row per thread (32x32b fragment), one sweep, weights go back to TMEM as the fp16 A operand of
      // the P.V contraction (see the file header).  st = this thread's fp32 state: everything drained from O so far plus
      // the chunks that took the exact path.
      q = 32 * lq + lane;
      Softmax2<C> st;
      st.init();
      float m_ref = -INFINITY, m_seen = -INFINITY;
      bool o_dirty = false;                               // warp uniform: P != 0 was stored since the last drain
      const uint32_t o_addr = tmem_base + g.o_col + 4 * wg + lane_addr;
      const int nvb = g.nvb, nck = g.G >> 1;
#ifdef CDS_PROFILE_SWITCHES
      unsigned cnt_chunks = 0, cnt_skipped = 0, cnt_exact = 0, cnt_drain = 0;
      long long ek_wait = 0, ek_work = 0, ph_ld = 0, ph_cls = 0, ph_w = 0, ph_st = 0, ph_tail = 0;
#endif
      // folds the O accumulator (reference m_ref) into st and clears it.  Every P.V issued so far must be complete.  There
      // is no commit per P.V (the MMA warp's per-tile work is a bubble on the tensor pipe); it issues P.V(T-2), S(T),
      // P.V(T-1), S(T+1) in this order and commits tfull after every S, so the accumulators of tile Tcur+1 being ready
      // implies that P.V(Tcur-1) and everything before it is done.  S(Tcur+1) does not depend on the epilogue of tile Tcur,
      // so waiting for it here cannot deadlock.  The last drain waits for the commit after the very last P.V instead.
      const uint32_t total_tiles = (uint32_t)(n_img * g.nchunks * g.nvb);
      auto drain = [&](uint32_t Tcur) {
        if (Tcur + 1u < total_tiles) mbar_wait(bar_tfull + 8 * ((Tcur + 1u) & 1u), ((Tcur + 1u) >> 1) & 1u, 9);
        else mbar_wait(bar_pvdone, 0u, 9);
        tc_fence_after();
        uint32_t o[4];
        tmem_ld4(o_addr, o);
        tmem_ld_wait4(o);
        const float mb = m_ref - PV_SHIFT, lb = __uint_as_float(o[3]);
        if (lb > 0.f) {
          const float M = fmaxf(st.m, mb);
          const float s0 = (st.m == -INFINITY) ? 0.f : ex2(st.m - M), s1 = ex2(mb - M);
          st.l = st.l * s0 + lb * s1;
#pragma unroll
          for (int c = 0; c < C; ++c) st.acc[c] = st.acc[c] * s0 + __uint_as_float(o[c]) * inv_scale * s1;
          st.m = M;
        }
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
        tmem_st4(o_addr, z);
        tmem_st_wait_all();
        o_dirty = false;
      };
      uint32_t T = 0;
      int unit = 0;
      for (int n = 0; n < n_img; ++n) {
        const float lw = __ldg(p.logw + n0 + n) * CDS_LOG2E;
        for (int ch = 0; ch < g.nchunks; ++ch, ++unit) {
          const int s = unit & 1;
          for (int vb = 0; vb < nvb; ++vb, ++T) {
            // re-base the reference before tile 1, 2, 4, 8, ... of the CTA's slice (tiles, not images: only the very first
            // tile runs without a reference, i.e. on the exact path -- a whole image would cost several per cent of a CTA
            // that owns ~70 images when the bank is sharded over 8 GPUs)
            if (T > 0u && (T & (T - 1u)) == 0u && T + 1u < total_tiles) {
              if (o_dirty) {
                drain(T);
#ifdef CDS_PROFILE_SWITCHES
                ++cnt_drain;
#endif
              }
              m_ref = m_seen;
            }
            const float off = lw - m_ref + PV_SHIFT;       // +inf while there is no reference yet (then every unit is exact)
            const float2 off2 = make_float2(off, off);
            const uint32_t buf = T & 1u;
#ifdef CDS_PROFILE_SWITCHES
            const long long ek0 = clock64();
#endif
            mbar_wait(bar_tfull + 8 * buf, (T >> 1) & 1u, 5);
#ifdef CDS_PROFILE_SWITCHES
            const long long ek1 = clock64();
            ek_wait += ek1 - ek0;
#endif
            tc_fence_after();
            const uint32_t taddr = tmem_base + buf * g.tmem_buf1 + lane_addr;
            // an 8-column block that runs past the end of the image row aliases the next row's granules; for the band's last
            // patch row those lie outside the staged norm-plane rows, so such columns carry no marker and could win the max
            // (and with it the reference): mask them explicitly.  Images whose width is a multiple of 8 never get here.
            const bool edge = 8 * vb + 8 > g.W;
            const int nval_v = g.Pw - 8 * vb, nval_u = g.Ph - g.chunk_u0[ch];
            // Work unit of a warp = 32 accumulator columns (two 16-column chunks = two P.V UMMAs) of its 32 rows; unit u belongs
            // to warpgroup u & 3.  Wide units matter: the epilogue is bound by instruction issue and latency, not by MUFU
            // (measured: moving exponentials to an FMA-pipe polynomial made it slower), and the per-unit overhead (tcgen05
            // ld / st with their warp syncs, votes, branches) is the same for 16 and 32 columns.
            // kind: 0 = every weight rounds to zero (store zeros), 1 = regular, 2 = exact path (overflow / no reference yet)
            auto unit32 = [&](int u, auto wide_tag) {
              constexpr int NC = decltype(wide_tag)::value;          // 32, or 16 for the odd last chunk of a tile
              const int j0c = 2 * u;                                 // first 16-column chunk of the unit
              uint32_t r[NC];
#ifdef CDS_PROFILE_SWITCHES
              const long long pk0 = clock64();
#endif
#if CDS_PV_ABLATE & 4            // A/B builds only: no tcgen05.ld (registers hold an arbitrary finite pattern)
              if (true) {
#pragma unroll
                for (int e = 0; e < NC; ++e) r[e] = __float_as_uint(-1000.f * (float)((e * 7 + lane) & 15));
              } else
#endif
              if constexpr (NC == 32) {
                tmem_ld_32x32b_x32(taddr + 16 * j0c, r);
                tmem_ld_wait32(r);
              } else {
                tmem_ld16(taddr + 16 * j0c, r);
                tmem_ld_wait16(r);
              }
#ifdef CDS_PROFILE_SWITCHES
              const long long pk1 = clock64();
              ph_ld += pk1 - pk0;
#endif
              if (edge) {
#pragma unroll
                for (int e = 0; e < NC; ++e)
                  if ((e & 7) >= nval_v || 2 * j0c + (e >> 3) >= nval_u) r[e] = 0xff800000u;
              }
              float mx = max3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
              float mx2 = max3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
#pragma unroll
              for (int e = 6; e + 3 < NC; e += 4) {
                mx = max3(mx, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
                mx2 = max3(mx2, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
              }
              mx = max3(mx, mx2, __uint_as_float(r[NC - 2]));
              mx = fmaxf(mx, __uint_as_float(r[NC - 1]));
              const float lg = fmaf(mx, c1, lw);           // best logit of the unit for this row (c1 > 0)
              m_seen = fmaxf(m_seen, lg);
              const float dd = lg - m_ref;
              int kind = 1;
              // every weight rounds to zero in fp16 (< 2^-33 of the reference, which never exceeds the best logit seen): up to
              // 4.5e6 such candidates add < 6e-4 of one top candidate's weight
              if (__all_sync(0xffffffffu, dd < -PV_SKIP)) kind = 0;
              // some weight would overflow fp16 (or there is no reference yet)
              else if (__any_sync(0xffffffffu, !(dd <= PV_EXCEED))) kind = 2;
#ifdef CDS_PROFILE_SWITCHES
              ++cnt_chunks;
              if (p.flags & 64) kind = 2;                  // profiling: 64 = every unit exact
#endif
              uint32_t pk[NC / 2];
#ifdef CDS_PROFILE_SWITCHES
              const long long pk2 = clock64();
              ph_cls += pk2 - pk1;
#endif
              if (kind == 1) {
#pragma unroll
                for (int e = 0; e < NC / 2; ++e) {
                  const float2 ar = fma2(make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1])), c1c1, off2);
#if (CDS_PV_ABLATE & 3) == 1      // A/B builds only: 1 = no MUFU, 2 = no F2FP, 3 = neither
                  pk[e] = pack_f16x2(ar.x * 1e-3f, ar.y * 1e-3f); continue;
#elif (CDS_PV_ABLATE & 3) == 2
                  pk[e] = __float_as_uint(ex2(ar.x)) ^ __float_as_uint(ex2(ar.y)); continue;
#elif (CDS_PV_ABLATE & 3) == 3
                  pk[e] = __float_as_uint(ar.x) ^ __float_as_uint(ar.y); continue;
#endif
                  if ((CDS_PV_POLY_MASK >> e) & 1) {
                    const float2 w = exp2_poly2(ar);
                    pk[e] = pack_f16x2(w.x, w.y);
                  } else {
                    pk[e] = pack_f16x2(ex2(ar.x), ex2(ar.y));
                  }
                }
                o_dirty = true;
              } else {
                if (kind == 2) {
                  // exact online softmax on the CUDA cores for the 32 rows x NC candidates of this warp, P = 0.  The V' blocks
                  // supply centre pixels and validity.
                  mbar_wait(bar_vready + 8 * s, (unit >> 1) & 1, 4);
                  const uint8_t* vblk = sStage + g.v_ring_off + (size_t)s * g.v_slot_bytes + 256 + (size_t)vb * g.v_tile_bytes +
                                        (size_t)(2 * j0c) * 128 + (size_t)(4 * (wg & 1)) * 16;
#pragma unroll
                  for (int e = 0; e < NC; ++e) {
                    const uint8_t* col = vblk + (e >> 3) * 128 + (e & 7) * 2;
                    if (*reinterpret_cast<const uint16_t*>(col + 3 * 16) != 0) {          // valid candidate (ones row)
                      float v[C];
#pragma unroll
                      for (int c = 0; c < C; ++c) v[c] = __half2float(*reinterpret_cast<const __half*>(col + c * 16)) * inv_scale;
                      st.push(fmaf(__uint_as_float(r[e]), c1, lw), v);
                    }
                  }
#ifdef CDS_PROFILE_SWITCHES
                  ++cnt_exact;
#endif
                }
#ifdef CDS_PROFILE_SWITCHES
                else ++cnt_skipped;
#endif
#pragma unroll
                for (int e = 0; e < NC / 2; ++e) pk[e] = 0u;
              }
              // P of chunk j sits in the first 8 columns of the chunk's 16
#if CDS_PV_ABLATE & 8            // A/B builds only: no tcgen05.st
              if (pk[0] != 0x12345u || pk[NC / 2 - 1] != 0x54321u) return;
#endif
#ifdef CDS_PROFILE_SWITCHES
              asm volatile("" : "+r"(pk[0]), "+r"(pk[NC / 2 - 1]));
              const long long pk3 = clock64();
              ph_w += pk3 - pk2;
#endif
              tmem_st8(taddr + 16 * j0c, pk);
              if constexpr (NC == 32) tmem_st8(taddr + 16 * (j0c + 1), pk + 8);
#ifdef CDS_PROFILE_SWITCHES
              ph_st += clock64() - pk3;
#endif
            };
#ifdef CDS_PROFILE_SWITCHES
            if (!(p.flags & 2))           // profiling: 2 = the epilogue touches nothing (tensor pipe + barriers only)
#endif
            for (int u = wg; 2 * u < nck; u += 4) {
              if (2 * u + 1 < nck) unit32(u, std::integral_constant<int, 32>{});
              else unit32(u, std::integral_constant<int, 16>{});
            }
#ifdef CDS_PROFILE_SWITCHES
            const long long ek2 = clock64();
#endif
            tmem_st_wait_all();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(bar_pready + 8 * buf);
#ifdef CDS_PROFILE_SWITCHES
            const long long ek3 = clock64();
            ek_work += ek3 - ek1;
            ph_tail += ek3 - ek2;
#endif
          }
        }
      }
      if (o_dirty) drain(T);
      m = st.m;
      l = st.l;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = st.acc[c];
#ifdef CDS_PROFILE_SWITCHES
      if (lane == 0) {
        atomicAdd(&g_els_counters[0], (unsigned long long)cnt_chunks);
        atomicAdd(&g_els_counters[1], (unsigned long long)cnt_skipped);
        atomicAdd(&g_els_counters[5], (unsigned long long)cnt_exact);
        atomicAdd(&g_els_counters[6], (unsigned long long)cnt_drain);
        if (warp == W_EPI0 + (int)(blockIdx.x & 15)) {           // one warp per CTA, a different one from CTA to CTA
          atomicAdd(&g_els_counters[11], (unsigned long long)ek_wait);
          atomicAdd(&g_els_counters[12], (unsigned long long)ek_work);
          atomicAdd(&g_els_counters[13], (unsigned long long)T);
          atomicAdd(&g_els_counters[14], (unsigned long long)ph_ld);
          atomicAdd(&g_els_counters[15], (unsigned long long)ph_cls);
          atomicAdd(&g_els_counters[2], (unsigned long long)ph_w);
          atomicAdd(&g_els_counters[3], (unsigned long long)ph_st);
          atomicAdd(&g_els_counters[4], (unsigned long long)ph_tail);
        }
      }
#endif
    } else {
    const int t4 = lane & 3, tr = lane >> 2;
    // the norm-plane marker suppresses invalid positions by 2^-(log2e * a^2/(2 beta) * 3 * INVALID_NORM); when beta -> 1 that
    // factor fades (a -> 0), so fall back to explicit column masking
    const bool weak_marker = CDS_LOG2E * a * a / (2.f * beta) * 3.f * INVALID_NORM < 64.f;
    // softmax state of row j = tr + 8*j over this thread's columns; even / odd columns accumulate separately (packed math)
    float m4[4];
    float2 l2[4], acc2[4][C];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      m4[j] = -INFINITY;
      l2[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) acc2[j][c] = make_float2(0.f, 0.f);
    }
    int want_dump = (p.dbg && split == 0) ? 1 : 0;
    uint32_t bar_t = bar_tfull;     // tfull at +0/+8, tempty at +16/+24
    asm volatile("" : "+r"(bar_t), "+r"(want_dump));   // opaque: one register each instead of per-tile recomputation
    const int nchunks = g.nchunks, nvb = g.nvb, vt_tile = g.vt_tile;
#ifdef CDS_PROFILE_SWITCHES
    const bool prof_pass1_only = (p.flags & 1) != 0, prof_mma_only = (p.flags & 2) != 0;   // CDS_DEBUG_FLAGS
#endif
#ifdef CDS_PROFILE_SWITCHES
    unsigned cnt_chunks = 0, cnt_skipped = 0, cnt_rescale = 0, cnt_tiles = 0, cnt_tiles_skipped = 0;
#endif
    uint32_t T = 0;
    int unit = 0;
    for (int n = 0; n < n_img; ++n) {
      const float lw = __ldg(p.logw + n0 + n) * CDS_LOG2E;
      const bool dump = want_dump != 0 && n == 0;
      for (int ch = 0; ch < nchunks; ++ch, ++unit) {
        const int s = unit & (S - 1);
        mbar_wait(bar_vready + 8 * s, (unit >> (S - 1)) & 1, 4);
        const float* vtile = reinterpret_cast<const float*>(sStage + (size_t)s * g.stage_bytes + g.vt_off);
        const int N = 8 * g.chunk_g[ch], u0 = g.chunk_u0[ch];
        for (int vb = 0; vb < nvb; ++vb, ++T, vtile += vt_tile) {
          const uint32_t buf = T & 1u;
          mbar_wait(bar_t + 8 * buf, (T >> 1) & 1u, 5);
          tc_fence_after();
          uint32_t taddr = tmem_base + buf * g.tmem_buf1 + lane_addr;
          int Nr = N;
          asm volatile("" : "+r"(taddr), "+r"(Nr));   // opaque: keep in registers instead of recomputing per chunk
          const float4* vt4 = reinterpret_cast<const float4*>(vtile) + t4;   // [(chunk*C + c)*4 + t4]
          // partial blocks inside the row and rounded-up patch rows are already masked by the norm plane's marker;
          // explicit masking is only needed when an 8-column block runs past the end of the image row
          const bool edge = weak_marker || 8 * vb + 8 > g.W;
          const int nval_v = g.Pw - 8 * vb, nval_u = g.Ph - u0;
#ifdef CDS_PROFILE_SWITCHES
          if (prof_mma_only) {     // profiling: MMA pipeline only
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            continue;
          }
#endif
          // r[0..7]: lanes +0..15, r[8..15]: lanes +16..31; element (row j, column pair h, lo) = r[EL(j,h) + lo],
          // its tile column = c0 + 8*h + 2*t4 + lo
#define EL(j, h) (8 * ((j) >> 1) + 2 * ((j) & 1) + 4 * (h))
          // `slow` = tiles that need explicit column masking or the debug dump; the common path carries neither check
          auto chunk = [&](uint32_t* r, int c0, const float4* pv, auto slow) {
            // columns that are not valid patches are forced to -inf for the max and get weight 0 explicitly in pass 2
            // (a thread may own nothing but masked columns: its state then stays (-inf, 0, 0) and drops out of the merge)
            bool mk[2][2] = {{false, false}, {false, false}};
            if (decltype(slow)::value && edge) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int lo = 0; lo < 2; ++lo) {
                  mk[h][lo] = 2 * t4 + lo >= nval_v || (c0 >> 3) + h >= nval_u;
                  if (mk[h][lo]) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[EL(j, h) + lo] = 0xff800000u;
                  }
                }
            }
            // pass 1: best logit per row of the chunk (c1 > 0, so the max commutes with the affine map)
            float cm[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float d = fmaxf(max3(__uint_as_float(r[EL(j, 0)]), __uint_as_float(r[EL(j, 0) + 1]),
                                         __uint_as_float(r[EL(j, 1)])), __uint_as_float(r[EL(j, 1) + 1]));
              cm[j] = fmaf(d, c1, lw);
            }
            if (decltype(slow)::value && dump) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int q = 32 * lq + tr + 8 * j, qi = i0 + (q >> 3), qj = j0 + (q & 7);
                if (qi < g.H && qj < g.W) {
                  float* dbg = p.dbg + ((size_t)b * g.H * g.W + (size_t)qi * g.W + qj) * ((size_t)g.Ph * g.Pw);
#pragma unroll
                  for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int lo = 0; lo < 2; ++lo) {
                      const int u = u0 + (c0 >> 3) + h, v = 8 * vb + 2 * t4 + lo;
                      if (u < g.Ph && v < g.Pw) dbg[u * g.Pw + v] = __uint_as_float(r[EL(j, h) + lo]) * inv_scale;
                    }
                }
              }
            }
            // every weight of this chunk is < 2^-40 of the running max for all rows of the warp: adding them
            // cannot change an fp32 sum (<= 4.5e6 candidates * 2^-40 = 4e-6 relative in the worst case)
            const float lead = fmaxf(max3(cm[0] - m4[0], cm[1] - m4[1], cm[2] - m4[2]), cm[3] - m4[3]);
#ifdef CDS_PROFILE_SWITCHES
            ++cnt_chunks;
            if (__all_sync(0xffffffffu, lead < -SKIP_LOG2)) { ++cnt_skipped; return; }
            if (__any_sync(0xffffffffu, lead > 0.f)) ++cnt_rescale;
#endif
            if (__all_sync(0xffffffffu, lead < -SKIP_LOG2)) return;
#ifdef CDS_PROFILE_SWITCHES
            if (prof_pass1_only) {     // profiling: pass 1 only
#pragma unroll
              for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m4[j], cm[j]);
              return;
            }
#endif
            if (lead > 0.f) {   // some row has a new maximum: rare after the first few images
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (cm[j] > m4[j]) {
                  const float sc = ex2(m4[j] - cm[j]);
                  const float2 sc2 = make_float2(sc, sc);
                  l2[j] = mul2(l2[j], sc2);
#pragma unroll
                  for (int c = 0; c < C; ++c) acc2[j][c] = mul2(acc2[j][c], sc2);
                  m4[j] = cm[j];
                }
            }
            // pass 2: weights and weighted sums
            float4 v[C];
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = pv[c * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float off = lw - m4[j];
              const float2 off2 = make_float2(off, off);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float2 ar = fma2(make_float2(__uint_as_float(r[EL(j, h)]), __uint_as_float(r[EL(j, h) + 1])), c1c1, off2);
                float2 w = make_float2(ex2(ar.x), ex2(ar.y));
                if (decltype(slow)::value) {
                  if (mk[h][0]) w.x = 0.f;
                  if (mk[h][1]) w.y = 0.f;
                }
                l2[j] = add2(l2[j], w);
#pragma unroll
                for (int c = 0; c < C; ++c)
                  acc2[j][c] = fma2(w, h == 0 ? make_float2(v[c].x, v[c].y) : make_float2(v[c].z, v[c].w), acc2[j][c]);
              }
            }
          };
#undef EL
          // no register double-buffering of the TMEM reads: four warpgroups at 96 registers hide the tcgen05.ld latency
          // better than three at 128 with a prefetch (171 vs 179 ms per trajectory step)
          auto ld = [&](int c0, uint32_t* r) {
            tmem_ld_16x256b_x2(taddr + c0, r);
            tmem_ld_16x256b_x2(taddr + (16u << 16) + c0, r + 8);
          };
          auto sweep = [&](auto slow) {
            uint32_t ra[16];
            // chunk i of tile T goes to warpgroup (i + T * chunks_per_tile) mod NUM_EPI_WG: the warpgroup that gets the
            // extra chunk of an uneven tile rotates, and the two TMEM buffers let the others run one tile ahead
            const int first = (int)((uint32_t)(wg + NUM_EPI_WG * 64 - (int)((T * (uint32_t)(Nr >> 4)) % NUM_EPI_WG)) % NUM_EPI_WG);
            const float4* pv = vt4 + first * (C * 4);
            for (int c0 = 16 * first; c0 < Nr; c0 += 16 * NUM_EPI_WG, pv += NUM_EPI_WG * C * 4) {
              ld(c0, ra);
              tmem_ld_wait16(ra);
              chunk(ra, c0, pv, slow);
            }
          };
#ifdef CDS_PROFILE_SWITCHES
          const unsigned c_before = cnt_chunks, s_before = cnt_skipped;
#endif
          if (edge || dump) sweep(std::true_type{});
          else sweep(std::false_type{});
#ifdef CDS_PROFILE_SWITCHES
          ++cnt_tiles;
          if (cnt_chunks - c_before == cnt_skipped - s_before) ++cnt_tiles_skipped;
#endif
          tc_fence_before();
          __syncwarp();
          if (elect_one()) mbar_arrive(bar_t + 16 + 8 * buf);
        }
        __syncwarp();
        if (elect_one()) mbar_arrive(bar_empty + 8 * s);
      }
    }
#ifdef CDS_PROFILE_SWITCHES
    if (lane == 0) {
      atomicAdd(&g_els_counters[0], (unsigned long long)cnt_chunks);
      atomicAdd(&g_els_counters[1], (unsigned long long)cnt_skipped);
      atomicAdd(&g_els_counters[2], (unsigned long long)cnt_rescale);
      atomicAdd(&g_els_counters[3], (unsigned long long)cnt_tiles);
      atomicAdd(&g_els_counters[4], (unsigned long long)cnt_tiles_skipped);
    }
#endif
    // the four threads t4 = 0..3 of a row group hold the same four rows over disjoint columns: merge them with shuffles,
    // after which thread (tr, t4) owns row tr + 8*t4 of its warp's 32
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float M = fmaxf(m4[j], __shfl_xor_sync(0xffffffffu, m4[j], 1));
      M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, 2));
      const float sc = (m4[j] == -INFINITY) ? 0.f : ex2(m4[j] - M);
      float lj = (l2[j].x + l2[j].y) * sc;
      lj += __shfl_xor_sync(0xffffffffu, lj, 1);
      lj += __shfl_xor_sync(0xffffffffu, lj, 2);
      if (t4 == j) { m = M; l = lj; }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float aj = (acc2[j][c].x + acc2[j][c].y) * sc;
        aj += __shfl_xor_sync(0xffffffffu, aj, 1);
        aj += __shfl_xor_sync(0xffffffffu, aj, 2);
        if (t4 == j) acc[c] = aj;
      }
    }
    q = 32 * lq + tr + 8 * t4;   // query row = TMEM lane
    }
    const int qi = i0 + (q >> 3), qj = j0 + (q & 7);
    // merge the warpgroups' partial softmax states and write this split's partials
    if (wg > 0) {
      float* dst = sMerge + ((wg - 1) * 128 + q) * (2 + C);
      dst[0] = m;
      dst[1] = l;
#pragma unroll
      for (int c = 0; c < C; ++c) dst[2 + c] = acc[c];
    }
    bar_sync_named(1, 128 * NUM_EPI_WG);
    if (wg == 0 && qi < g.H && qj < g.W) {
      float M = m;
#pragma unroll
      for (int w = 1; w < NUM_EPI_WG; ++w) M = fmaxf(M, sMerge[((w - 1) * 128 + q) * (2 + C)]);
      const float w0 = (m == -INFINITY) ? 0.f : ex2(m - M);
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
  if (warp == W_B0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// per-k "norm plane": granule (n,u,x) = three-way fp16 split of |p(n,u,x)|^2 arranged (ph,pm,pl,ph,pm,ph,0,0);
// positions that are not the top-left corner of a valid k x k patch carry INVALID_NORM
__global__ void norm_plane_kernel(const float* __restrict__ images, long long N, int C, int H, int W, int k,
                                  uint4* __restrict__ out) {
  const long long total = N * H * W;
  const int Ph = H - k + 1, Pw = W - k + 1;
  for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < total;
       gidx += (long long)gridDim.x * blockDim.x) {
    const int v = gidx % W, u = (gidx / W) % H;
    const long long n = gidx / ((long long)W * H);
    __half h[8];
    const __half z = __float2half_rn(0.f);
    if (u < Ph && v < Pw) {
      const float* img = images + n * C * H * W;
      float s = 0.f;
      for (int c = 0; c < C; ++c)
        for (int dy = 0; dy < k; ++dy)
          for (int dx = 0; dx < k; ++dx) {
            const float t = img[(c * H + u + dy) * W + v + dx];
            s = fmaf(t, t, s);
          }
      const __half ph = __float2half_rn(s);
      const __half pm = __float2half_rn(s - __half2float(ph));
      const __half pl = __float2half_rn(s - __half2float(ph) - __half2float(pm));
      h[0] = ph; h[1] = pm; h[2] = pl; h[3] = ph; h[4] = pm; h[5] = ph; h[6] = z; h[7] = z;
    } else {
      const __half big = __float2half_rn(INVALID_NORM);
      h[0] = big; h[1] = big; h[2] = big; h[3] = z; h[4] = z; h[5] = z; h[6] = z; h[7] = z;
    }
    out[gidx] = *reinterpret_cast<uint4*>(h);
  }
}

}  // namespace

#ifdef CDS_PROFILE_SWITCHES
// profile builds only (not part of include/cdscore.h): copies the counters to the host (synchronises) and clears them
extern "C" int cds_debug_els_counters(unsigned long long* out_host8) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out_host8, g_els_counters, sizeof(g_els_counters));
  unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_els_counters, z, sizeof(z));
  return e == cudaSuccess ? CDS_OK : CDS_ERR_CUDA;
}
#endif

extern "C" int64_t cds_els_umma_smem_bytes(int C, int H, int W, int k, int passes, int bank_planes) {
  UmmaGeom g;
  uint2 local[MAX_MMAS];
  return make_geom(C, H, W, k, passes, bank_planes, g, local) ? g.smem_total : 0;
}

extern "C" int cds_els_umma_pv_supported(int C, int H, int W, int k, int passes, int bank_planes) {
  UmmaGeom g;
  uint2 local[MAX_MMAS];
  return make_geom(C, H, W, k, passes, bank_planes, g, local, 1) ? 1 : 0;
}

extern "C" int cds_pack_norm_plane(const float* images, int64_t N, int C, int H, int W, int k, void* out_f16,
                                   void* stream) {
  CDS_CHECK_ARG(N >= 1 && k >= 1 && k <= H && k <= W, "cds_pack_norm_plane: bad arguments");
  const long long total = (long long)N * H * W;
  const int threads = 256;
  const int blocks = (int)((total + threads - 1) / threads > 148 * 32 ? 148 * 32 : (total + threads - 1) / threads);
  norm_plane_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(images, N, C, H, W, k, (uint4*)out_f16);
  CDS_CHECK_LAUNCH("norm_plane_kernel");
  return CDS_OK;
}

namespace {
// geometry for one epilogue variant: prefers the mixed K layout (rows8 plane) when it keeps the tiling and saves >= 8 %
// of the UMMAs.  Measured on 32x32x3: k=17 8.5 -> 6.5 ms; k=9 with the band halved to fit shared memory was 12 % slower
// although it issues 39 % fewer UMMAs.
bool pick_geom(int C, int H, int W, int k, int passes, int planes, int pv, bool try_mixed, UmmaParams& p) {
  bool have = false;
  if (try_mixed) {
    UmmaGeom gv;
    static thread_local uint2 scratch[MAX_MMAS];
    const bool okv = make_geom(C, H, W, k, passes, planes, gv, scratch, pv) != 0;
    const bool okm = make_geom(C, H, W, k, passes, planes, p.g, p.table, pv, 1) != 0;
    have = okm && (!okv || (p.g.stages >= gv.stages && p.g.G == gv.G && p.g.n_mma * 100 < gv.n_mma * 92));
  }
  return have || make_geom(C, H, W, k, passes, planes, p.g, p.table, pv) != 0;
}
}  // namespace

extern "C" int cds_els_partials_umma_window(int query_pad, const float* x, int B, int C, int H, int W, int k,
                                            const float* beta, const void* bank_hi, const void* bank_lo,
                                            const void* bank_rows, float bank_scale, const void* norm_plane,
                                            const int32_t* idx, const float* logw, int64_t n_sel, int splits, int passes,
                                            int variant, int qi0, int qj0, int qrows, int qcols, float* m, float* l,
                                            float* acc, float* dbg_dots, void* stream) {
  UmmaParams p;
  CDS_CHECK_ARG(qi0 >= 0 && qj0 >= 0 && qrows >= 1 && qcols >= 1 && qi0 + qrows <= H && qj0 + qcols <= W,
                "cds_els_partials_umma_window: window (%d,%d)+(%d,%d) outside the %dx%d image", qi0, qj0, qrows, qcols, H, W);
  p.qi0 = qi0; p.qj0 = qj0; p.qrows = qrows; p.qcols = qcols;
  const int planes = bank_lo ? 2 : 1;
  const char* mx = getenv("CDS_ELS_MIXED");     // A/B switch: 0 = vertical granules only
  const bool try_mixed = bank_rows != nullptr && !(mx && atoi(mx) == 0);
  CDS_CHECK_ARG(variant >= CDS_ELS_AUTO && variant <= CDS_ELS_PV, "cds_els_partials_umma: unknown variant %d", variant);
  // P.V epilogue: single-plane banks, no dot-product dump.  "auto" uses it where it was measured faster (profiles/r02*):
  // up to k = 9 the kernel is bound by the epilogue; from k = 11 by the main contraction, where the extra N=16 UMMAs of the
  // P.V contraction (~29 cycles each) cost more than the FMA-pipe weighted sum they replace.
  bool pv = false;
  if (variant != CDS_ELS_FMA && dbg_dots == nullptr) {
    const char* pk = getenv("CDS_PV_MAX_K");    // A/B switch for the auto rule
    const int pv_max_k = pk ? atoi(pk) : 9;
    if (variant == CDS_ELS_PV || k <= pv_max_k) pv = pick_geom(C, H, W, k, passes, planes, 1, try_mixed, p);
  }
  if (variant == CDS_ELS_PV && !pv) {
    cds_set_error("cds_els_partials_umma: the P.V epilogue does not support C=%d H=%d W=%d k=%d passes=%d planes=%d%s", C, H,
                  W, k, passes, planes, dbg_dots ? " with a dot-product dump" : "");
    return CDS_ERR_UNSUPPORTED;
  }
  if (!pv && !pick_geom(C, H, W, k, passes, planes, 0, try_mixed, p)) {
    cds_set_error("cds_els_partials_umma: unsupported geometry C=%d H=%d W=%d k=%d passes=%d planes=%d", C, H, W, k,
                  passes, planes);
    return CDS_ERR_UNSUPPORTED;
  }
  if (getenv("CDS_DEBUG_GEOM"))
    fprintf(stderr, "cdscore: els_umma k=%d passes=%d planes=%d pv=%d mixed=%d G=%d chunks=%d nvb=%d n_mma=%d n_tmem=%d stages=%d smem=%d\n",
            k, passes, planes, pv ? 1 : 0, p.g.rem ? 1 : 0, p.g.G, p.g.nchunks, p.g.nvb, p.g.n_mma, p.g.n_tmem, p.g.stages, p.g.smem_total);
  CDS_CHECK_ARG(B >= 1 && n_sel >= 1 && splits >= 1, "cds_els_partials_umma: empty problem");
  if (splits > n_sel) splits = (int)n_sel;
  p.B = B; p.pad = query_pad; p.splits = splits; p.n_sel = n_sel;
  p.x = x; p.beta = beta;
  p.bank_hi = (const uint8_t*)bank_hi; p.bank_lo = (const uint8_t*)bank_lo;
  p.bank_rows = (const uint8_t*)bank_rows;
  p.norm_plane = (const uint8_t*)norm_plane;
  p.scale = bank_scale;
  p.idx = idx; p.logw = logw;
  p.m = m; p.l = l; p.acc = acc; p.dbg = dbg_dots;
  make_pv_table(p.g, p.pv_table);
  {
    const char* f = getenv("CDS_DEBUG_FLAGS");
    p.flags = f ? atoi(f) : 0;
    const char* at = getenv("CDS_A_TMEM");     // A/B switch: 0 = every query slice from shared memory
    if (at && atoi(at) == 0) p.g.n_tmem = p.g.n_h;   // the horizontal slices of the mixed layout exist only in TMEM
  }
  const int tiles = ((qrows + TI - 1) / TI) * ((qcols + TJ - 1) / TJ);
  dim3 grid(tiles, splits, B);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define LAUNCH(CC)                                                                                                        \
  case CC:                                                                                                                \
    if (pv) {                                                                                                             \
      e = cudaFuncSetAttribute(els_umma_kernel<CC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g.smem_total);   \
      if (e == cudaSuccess) els_umma_kernel<CC, true><<<grid, THREADS, p.g.smem_total, st>>>(p);                          \
    } else {                                                                                                              \
      e = cudaFuncSetAttribute(els_umma_kernel<CC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.g.smem_total);  \
      if (e == cudaSuccess) els_umma_kernel<CC, false><<<grid, THREADS, p.g.smem_total, st>>>(p);                         \
    }                                                                                                                     \
    break;
  switch (C) {
    LAUNCH(1) LAUNCH(2) LAUNCH(3)
  }
#undef LAUNCH
  if (e != cudaSuccess) {
    cds_set_error("els_umma_kernel attribute: %s", cudaGetErrorString(e));
    return CDS_ERR_CUDA;
  }
  CDS_CHECK_LAUNCH("els_umma_kernel");
  return CDS_OK;
}

extern "C" int cds_els_partials_umma(int query_pad, const float* x, int B, int C, int H, int W, int k, const float* beta,
                                     const void* bank_hi, const void* bank_lo, const void* bank_rows, float bank_scale,
                                     const void* norm_plane, const int32_t* idx, const float* logw, int64_t n_sel,
                                     int splits, int passes, int variant, float* m, float* l, float* acc, float* dbg_dots,
                                     void* stream) {
  return cds_els_partials_umma_window(query_pad, x, B, C, H, W, k, beta, bank_hi, bank_lo, bank_rows, bank_scale, norm_plane,
                                      idx, logw, n_sel, splits, passes, variant, 0, 0, H, W, m, l, acc, dbg_dots, stream);
}
