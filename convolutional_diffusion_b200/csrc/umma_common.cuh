// Shared pieces of the tcgen05 ELS kernels: geometry, launch parameters, PTX wrappers, descriptor-table builder.
#pragma once
#include "common.cuh"
#include "../../include/cdscore.h"

namespace umma {


constexpr int TI = 16, TJ = 8;           // query tile: 16 rows x 8 columns = 128 queries = UMMA M
constexpr int MAX_STAGES = 2;
constexpr int MAX_CHUNKS = 8;
#ifndef CDS_NUM_EPI_WG
#define CDS_NUM_EPI_WG 4
#endif
constexpr int NUM_EPI_WG = CDS_NUM_EPI_WG;            // epilogue warpgroups: all of them consume every tile, each a quarter of its columns
constexpr int THREADS = 128 + 128 * NUM_EPI_WG;   // 4 control warps + the epilogue warpgroups
constexpr int TMEM_COLS = 512;
constexpr int MAX_MMAS = 320;
constexpr float SKIP_LOG2 = 40.f;        // chunks whose weights are all < 2^-40 of the running max are skipped
constexpr float INVALID_NORM = 60000.f;  // norm-plane marker of positions that are not valid patches
// P.V epilogue (weights as an fp16 tensor-core operand): P = 2^(logit - m_ref + PV_SHIFT)
constexpr float PV_SHIFT = 8.f;          // weights down to 2^-22 of the reference stay normal fp16 numbers, down to 2^-33 nonzero
constexpr float PV_EXCEED = 7.5f;        // a logit more than this above the reference would overflow fp16 -> exact SIMT path
constexpr float PV_SKIP = 33.f;          // every logit of a chunk more than this below the reference: P rounds to 0 anyway

struct UmmaGeom {
  int C, H, W, k, d, Ph, Pw;
  int nb;             // dy blocks of 8 rows
  int RA;             // bytes between query rows in the A tile = (k+7)*16
  int a_block;        // bytes of one (c,b) block of A = TI*RA
  int a_plane;        // bytes of one precision plane of A = C*nb*a_block
  int a_const;        // byte offset of the constant (-a*scale/2) block in A
  int a_zero;         // byte offset of the zero block in A
  int a_bytes;
  int S1;             // bytes of one image row of granules = W*16
  int G, R;           // patch rows per band (N = 8*G), strip rows staged per band and channel
  int chan_bytes;     // H*W*16: one channel of one image (and one norm plane) in HBM
  int img_bytes;      // C*R*W*16: one precision plane of a band in shared memory
  int np_bytes;       // G*W*16: the band's rows of the norm plane
  int tile_pad;       // zeroed guard after each staged tile
  int np_off;         // offset of the norm plane inside a stage
  int vt_off;         // offset of the centre-pixel table inside a stage
  int vt_tile;        // floats per tile of that table: per 16-column chunk and channel 16 floats in fragment order; PV variant: the
                      // fp32 (tf32) UMMA operand V'^T [16 rows][N] K-major = N*64 bytes per tile
  int pv;             // 1 = geometry of the P.V (weighted sum on the tensor cores) variant
  int o_col;          // P.V: TMEM column of the O accumulator (16 columns: 4 per epilogue warpgroup = C sums + denominator)
  // P.V: the V' operand lives in its own two-slot ring behind the stages (slot = band & 1), so that a stage can go back
  // to the producer as soon as the main UMMAs of its band are done, while V' stays until the band's last P.V (one tile
  // later).  Slot layout: per tile [256 zero bytes][per patch row one 128-byte core matrix: 8 rows x 8 candidates fp16], and
  // 256 more zero bytes behind the last tile; the zero block of tile vb+1 is the trailing zero block of tile vb.
  int v_ring_off;     // offset of slot 0 from the first stage
  int v_slot_bytes;
  int v_tile_bytes;   // distance between tiles = 256 + G * 128; the operand blocks of a tile start 256 bytes into it
  int stage_bytes;
  int stages;
  int nchunks, chunk_u0[MAX_CHUNKS], chunk_g[MAX_CHUNKS];
  int nvb;            // 8-column blocks of candidate patches
  int n_mma;          // descriptor-table entries per accumulator tile (all precision combinations)
  int rem, nbh;       // mixed K layout: the last rem = k - 8*nb patch rows are contracted as nbh horizontal 8-pixel granules
                      // per row ("rows8" bank plane) instead of one more mostly-empty vertical block; rem = 0: off
  int h_plane;        // the horizontal query regions are TMEM resident; they are built once in the (not yet used) staging
                      // area: region (plane,c,r,blk) = TI*128 bytes at plane*h_plane + ..., followed by TI*128 zero bytes
  int n_h;            // table entries (always the first ones) that contract horizontal granules
  int Rh;             // rows of the rows8 band staged per channel = G + rem - 1
  int hb_off;         // offset of that band inside a stage
  int tmem_buf1;      // TMEM column of the second accumulator buffer (the first sits at column 0)
  int a_tmem_col;     // TMEM column of the resident query K slices (8 columns = one K=16 slice of all 128 rows)
  int n_tmem;         // the first n_tmem table entries take their A operand from TMEM instead of shared memory
  int passes, bank_planes;
  int smem_A, smem_stage, smem_merge, smem_table, smem_bar, smem_total;
};

struct UmmaParams {
  UmmaGeom g;
  int B, pad, splits;
  int qi0, qj0, qrows, qcols;   // query window in pixels: the whole image, or the bbELS centre region
  long long n_sel;
  const float* x;
  const float* beta;
  const uint8_t* bank_hi;
  const uint8_t* bank_lo;
  const uint8_t* bank_rows;   // rows8 plane (horizontal granules) or null
  const uint8_t* norm_plane;
  float scale;
  const int32_t* idx;
  const float* logw;
  float *m, *l, *acc, *dbg;
  int flags;               // profiling switches, honoured only when built with -DCDS_PROFILE_SWITCHES (CDS_DEBUG_FLAGS):
                           // 1 = pass 1 only, 2 = UMMAs only (no epilogue math), 8 = epilogue only (no UMMAs issued)
  uint2 table[MAX_MMAS];   // lo words of the (A,B) descriptors relative to the A base / the tile origin in a stage
  uint2 pv_table[16];      // P.V: per 16-column chunk the V' descriptor (lo, hi word) relative to the tile's origin in its slot
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must become a trap (CUDA error), never a hung GPU.  2^30 polls: a try_wait that fails
// suspends the thread for a hardware-chosen slice first, so this is minutes of wall clock -- far beyond any healthy wait,
// also under a profiler or time-slicing (round 1 trapped after 2^22 polls, which ncu replays could reach).  A time-based
// bound (%globaltimer) was tried and dropped: inline it costs the FMA epilogue registers (spills 16 -> 28 bytes), out of
// line the call spills more.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  for (uint32_t it = 0; it < (1u << 30); ++it)
    if (mbar_try(bar, parity)) return;
  printf("cdscore: mbarrier timeout tag=%d block=(%d,%d,%d) thread=%d\n", tag, blockIdx.x, blockIdx.y, blockIdx.z,
         threadIdx.x);
  __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
}
// same with the A operand resident in TMEM (lane = row, one 32-bit column = two consecutive K elements)
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
__device__ __forceinline__ void tmem_st_wait_all() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 16 columns: r0,r1 = (lane t/4, columns 2*(t%4)+{0,1}), r2,r3 = (lane t/4+8, same columns), r4..r7 = the
// same rows, columns +8 (cute SM100_TMEM_LOAD_16dp256b2x fragment layout)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// the wait names the destination registers as in/out operands so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for a 32-register load: the second statement ties registers 16..31 to the wait (volatile statements keep their order)
__device__ __forceinline__ void tmem_ld_wait32(uint32_t* r) {
  tmem_ld_wait16(r);
  asm volatile(""
               : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait4(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// two fp32 -> one packed f16x2 register (lo = first argument = the lower K index of a TMEM-resident A operand)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// packed fp32x2 arithmetic (sm_100+): two FMAs per issue slot
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for two values on the FMA pipe (no MUFU): round-to-nearest split x = n + f with the 1.5*2^23 trick, degree-3 minimax
// polynomial for 2^f on [-0.5, 0.5] (max relative error 7.6e-5, a third of the fp16 half-ulp of the weights it feeds),
// exponent patched in with an integer add.  x is clamped at -30 (the result, < 2^-30 * scale, rounds to 0 in fp16 anyway);
// the caller guarantees x < 100.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
  const float2 m1 = make_float2(-1.f, -1.f);
  x.x = fmaxf(x.x, -30.f);
  x.y = fmaxf(x.y, -30.f);
  const float2 t = add2(x, magic);
  const float2 n = add2(t, nmagic);
  const float2 f = fma2(n, m1, x);
  float2 q = fma2(f, make_float2(0.05520550534129143f, 0.05520550534129143f), make_float2(0.24261397123336792f, 0.24261397123336792f));
  q = fma2(q, f, make_float2(0.6932547688484192f, 0.6932547688484192f));
  q = fma2(q, f, make_float2(0.9999276995658875f, 0.9999276995658875f));
  return make_float2(__uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23)),
                     __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23)));
}
// one lane of a converged warp (the warp stays converged around it, so descriptors live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major, no-swizzle shared memory descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4   [16,30) LBO>>4 (stride between the two K granules)   [32,46) SBO>>4 (stride between
//   8-row groups)   [46,48) version = 1   [61,64) layout = 0 (interleave / no swizzle)
__device__ __forceinline__ uint64_t desc_hi(uint32_t sbo_bytes) {
  return ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}



// ------------------------------------------------------------------ host side: geometry + descriptor table
struct Gran { int a, b; };

// pairs consecutive K granules of one precision combination into UMMA descriptors (K = 16 = two granules);
// an odd tail is paired with the zero block on the query side
inline int emit_pairs(const Gran* gr, int n, int a_zero, uint2* table, int nm) {
  for (int q = 0; q < n; q += 2) {
    if (nm >= MAX_MMAS) return -1;
    int la, lb;
    if (q + 1 < n) { la = gr[q + 1].a - gr[q].a; lb = gr[q + 1].b - gr[q].b; }
    else           { la = a_zero - gr[q].a;      lb = 16; }
    if (la <= 0 || lb <= 0 || (la >> 4) > 0x3FFF || (lb >> 4) > 0x3FFF) return -1;
    table[nm].x = (uint32_t)(gr[q].a >> 4) | ((uint32_t)(la >> 4) << 16);
    table[nm].y = (uint32_t)(gr[q].b >> 4) | ((uint32_t)(lb >> 4) << 16);
    ++nm;
  }
  return nm;
}

// mixed = 1: contract the patch rows beyond the last full block of 8 through the rows8 plane (needs k > 8, k % 8 != 0,
// a single-plane bank); returns 0 when that layout does not apply or does not fit
inline int make_geom(int C, int H, int W, int k, int passes, int bank_planes, UmmaGeom& g, uint2* table, int pv = 0,
                     int mixed = 0) {
  if (C < 1 || C > 3 || H > 64 || W > 64 || H < k || W < k || (k & 1) == 0 || k < 3) return 0;
  if (passes < 1 || passes > 2 || bank_planes < 1 || bank_planes > 2) return 0;
  if (mixed && (bank_planes > 1 || k < 9 || (k & 7) == 0)) return 0;
  if (pv && bank_planes > 1) return 0;          // the value operand of the P.V contraction is the single exact fp16 plane
  g.C = C; g.H = H; g.W = W; g.k = k; g.d = k / 2;
  g.Ph = H - k + 1; g.Pw = W - k + 1;
  g.nb = mixed ? k / 8 : (k + 7) / 8;
  g.rem = mixed ? k - 8 * g.nb : 0;
  g.nbh = mixed ? (k + 7) / 8 : 0;
  g.RA = (k + 7) * 16;
  g.a_block = TI * g.RA;
  g.a_plane = C * g.nb * g.a_block;
  g.h_plane = C * g.rem * g.nbh * TI * 128;
  g.a_const = passes * g.a_plane;
  g.a_zero = g.a_const + g.a_block;
  g.a_bytes = g.a_zero + g.a_block;
  g.S1 = W * 16;
  g.chan_bytes = H * W * 16;
  g.tile_pad = ((k + 8) * 16 + 127) / 128 * 128;
  g.passes = passes; g.bank_planes = bank_planes;
  g.nvb = (g.Pw + 7) / 8;
  g.pv = pv;
  g.smem_A = (g.a_bytes + 1023) / 1024 * 1024;
  // the merge scratch of the warpgroups' final states ([NUM_EPI_WG-1][128][2+C] floats) aliases the query tile: by then
  // every UMMA that reads the tile has completed (each epilogue warp has waited for the last accumulator tile)
  g.smem_merge = 0;
  if (g.a_bytes < (NUM_EPI_WG - 1) * 128 * (2 + C) * 4) return 0;
  g.smem_bar = 8 * 16 + 16 + 32;
  const int n_gran = C * g.nb * k;
  const int nm_max = (passes + bank_planes - 1) * ((n_gran + 2) / 2 + (C * g.rem * g.nbh + 1) / 2);
  if (nm_max > MAX_MMAS) return 0;
  g.smem_table = 128;
  const int fixed = g.smem_A + g.smem_merge + g.smem_table + g.smem_bar + 1024;

  // Staging unit = (image, band of G patch rows): the band holds the strip rows u0 .. u0+R-1 of every channel,
  // R = G + max(8*(nb-1), d) (dy blocks of the last patch row; the centre pixels of the band's patches), plus G rows of
  // the norm plane and the centre-pixel tables of its nvb tiles.  N = 8*G <= 256 candidates per UMMA, G even
  // (UMMA M=128 needs N % 16 == 0).  Pick the largest G that leaves room for two stages.
  const int halo = 8 * (g.nb - 1) > g.d ? 8 * (g.nb - 1) : g.d;
  int bestG = 0, bestStages = 0, bestWaste = 1 << 30;
  for (int G = 32; G >= 2; G -= 2) {
    if (G > ((g.Ph + 1) & ~1)) continue;
    const int nch = (g.Ph + G - 1) / G;
    if (nch > MAX_CHUNKS || nch * G > H) continue;          // norm-plane / strip rows of a partial last band must exist
    if (pv && 16 * G + 16 > TMEM_COLS) continue;            // P.V variant: two S buffers + the 16-column O accumulator
    const int R = G + halo;
    const int band = C * R * g.S1;
    const int hband = g.rem ? C * (G + g.rem - 1) * g.S1 + g.tile_pad : 0;
    const int vt_tile = pv ? 0 : 8 * G / 2 * 6;            // floats per tile of the centre-pixel table (FMA epilogue)
    const int vring = pv ? 2 * (g.nvb * (256 + G * 128) + 256) : 0;
    const int stage = (bank_planes * (band + g.tile_pad) + hband + G * g.S1 + g.tile_pad + g.nvb * vt_tile * 4 + 127) / 128 * 128;
    const int st = fixed + vring + 2 * stage <= 227 * 1024 ? 2 : (fixed + vring + stage <= 227 * 1024 ? 1 : 0);
    // prefer two stages, then the cheapest tiling: candidate columns per image incl. padded patch rows, plus a fixed
    // per-tile overhead worth ~32 columns (barrier round trips, pipeline fill)
    const int waste = nch * (8 * G + 32);
    if (st > bestStages || (st == bestStages && st > 0 && waste < bestWaste)) { bestStages = st; bestG = G; bestWaste = waste; }
  }
  if (bestG == 0) return 0;
  if (pv && bestStages < 2) return 0;                        // the P.V pipeline releases a stage two tiles late
  g.G = bestG; g.R = bestG + halo;
  g.Rh = g.rem ? bestG + g.rem - 1 : 0;
  g.nchunks = (g.Ph + g.G - 1) / g.G;
  for (int c = 0; c < g.nchunks; ++c) { g.chunk_u0[c] = c * g.G; g.chunk_g[c] = g.G; }
  g.img_bytes = C * g.R * g.S1;                              // one precision plane of a band in shared memory
  g.np_bytes = g.G * g.S1;
  g.hb_off = bank_planes * (g.img_bytes + g.tile_pad);
  g.np_off = g.hb_off + (g.rem ? C * g.Rh * g.S1 + g.tile_pad : 0);
  g.vt_off = g.np_off + g.np_bytes + g.tile_pad;
  g.vt_tile = pv ? 0 : 8 * g.G / 2 * 6;                     // floats per tile (see UmmaGeom::vt_tile)
  g.v_tile_bytes = 256 + g.G * 128;
  g.v_slot_bytes = pv ? g.nvb * g.v_tile_bytes + 256 : 0;
  g.stage_bytes = (g.vt_off + g.nvb * g.vt_tile * 4 + 127) / 128 * 128;
  g.stages = bestStages;

  // K granule lists.  Mixed layout first: the horizontal granules of every query plane (table.x bit 31: the A operand
  // of these entries exists only in TMEM, built from a scratch copy with 128-byte row groups), then per precision
  // combination (query plane, bank plane): (0,0)+norm granule [, (1,0)] [, (0,1)]
  Gran gr[3 * 4 * 32 + 2];
  int nm = 0;
  if (g.rem) {
    for (int pa = 0; pa < passes; ++pa) {
      int n = 0;
      for (int c = 0; c < C; ++c)
        for (int r = 0; r < g.rem; ++r)
          for (int blk = 0; blk < g.nbh; ++blk) {
            gr[n].a = pa * g.h_plane + ((c * g.rem + r) * g.nbh + blk) * TI * 128;
            gr[n].b = g.hb_off + ((c * g.Rh + r) * W + 8 * blk) * 16;
            ++n;
          }
      nm = emit_pairs(gr, n, passes * g.h_plane, table, nm);
      if (nm < 0) return 0;
    }
    for (int t = 0; t < nm; ++t) table[t].x |= 0x80000000u;
    if (passes * g.h_plane + TI * 128 > g.stages * g.stage_bytes) return 0;
  }
  g.n_h = nm;
  for (int cb = 0; cb < passes + bank_planes - 1; ++cb) {
    const int pa = (cb == 1 && passes > 1) ? 1 : 0;
    const int pb = (cb > 0 && !pa) ? 1 : 0;
    int n = 0;
    for (int c = 0; c < C; ++c)
      for (int blk = 0; blk < g.nb; ++blk)
        for (int dx = 0; dx < k; ++dx) {
          gr[n].a = pa * g.a_plane + (c * g.nb + blk) * g.a_block + dx * 16;
          gr[n].b = pb * (g.img_bytes + g.tile_pad) + (c * g.R + 8 * blk) * g.S1 + dx * 16;   // band-relative rows
          ++n;
        }
    if (cb == 0) { gr[n].a = g.a_const; gr[n].b = g.np_off; ++n; }
    nm = emit_pairs(gr, n, g.a_zero, table, nm);
    if (nm < 0) return 0;
  }
  g.n_mma = nm;
  // TMEM: two accumulator buffers of 8*G columns; what is left holds query K slices, which the tensor core then reads
  // from TMEM instead of re-reading them from shared memory for every candidate tile
  g.tmem_buf1 = 8 * g.G;
  g.o_col = TMEM_COLS - 16;                                  // P.V: the O accumulator takes the last 16 columns
  g.a_tmem_col = 2 * g.tmem_buf1;
  g.n_tmem = (TMEM_COLS - (pv ? 16 : 0) - g.a_tmem_col) / 8;
  if (g.n_tmem > nm) g.n_tmem = nm;
  if (g.n_tmem < g.n_h) return 0;                            // the horizontal slices have no shared-memory copy
  g.v_ring_off = g.stages * g.stage_bytes;
  g.smem_stage = g.stages * g.stage_bytes + 2 * g.v_slot_bytes;
  g.smem_total = fixed + g.smem_stage;
  return g.smem_total <= 227 * 1024;
}


// V' descriptors of the P.V contraction (K-major, no swizzle, LBO = 128 B between the two K granules): chunk j = patch rows
// 2j, 2j+1 of the tile = two 128-byte core matrices at 256 + 256*j.  Its owner warpgroup w = (j >> 1) & 3 accumulates in O columns
// 4w..4w+3: for w < 2 the value rows are the first 8-row group of the operand and the second group is the zero block behind
// the tile; for w >= 2 the first group is the zero block in front of the tile and the value rows are the second group
// (SBO = distance between the groups).
inline void make_pv_table(const UmmaGeom& g, uint2* t) {
  const int nck = g.G / 2;
  for (int j = 0; j < 16; ++j) {
    const int data = 256 + 256 * (j < nck ? j : 0), zpost = 256 + g.G * 128;
    const int w = (j >> 1) & 3;              // the epilogue works in units of two chunks (32 columns); unit u belongs to warpgroup u & 3
    const int start = w < 2 ? data : 0, sbo = w < 2 ? zpost - data : data;
    t[j].x = (uint32_t)(start >> 4) | (8u << 16);
    t[j].y = (uint32_t)(sbo >> 4) | (1u << 14);          // bit 46 of the descriptor: version 1
  }
}

}  // namespace umma
