// Microbenchmarks behind the epilogue design of csrc/els_umma.cu (not part of the product library):
//   ld    : tcgen05.ld throughput per SM for the fragment shapes the epilogue uses, 4..16 reader warps
//   st    : tcgen05.st throughput
//   mufu  : ex2.approx rate per SM (the other 8-cycles-per-column floor)
//   umma  : dispatch cost of small tcgen05.mma tiles (M=128, N=16..240, K=16, kind::f16; A from TMEM or shared
//           memory), i.e. what a P.V contraction on the tensor pipe costs next to the main contraction
// Build + run (on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tmem_bench tests/microbench/tmem_bench.cu
//   gpurun_out/tmem_bench
// Every number is cycles of the SM clock (clock64) inside one CTA per SM, max over the 148 CTAs.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../convolutional_diffusion_b200/csrc/umma_common.cuh"

// The product's error plumbing is not linked here.
void cds_set_error(const char*, ...) {}

using namespace umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
struct Res { unsigned long long cycles; unsigned sink; };

// mode 0: 32x32b.x16   1: 32x32b.x32   2: 16x256b.x2 on both lane halves (the product's fragment)   3: st 32x32b.x16
// 4: ex2 only (16 per iteration)   5: ld 32x32b.x16 with two loads in flight before each wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) ldst_kernel(int iters, Res* out) {
  __shared__ uint32_t s_tmem;
  __shared__ unsigned long long s_t0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = s_tmem + (((uint32_t)((warp & 3) * 32)) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 31 + i;
  // make the columns finite before they are read
  for (int c = 0; c < 512; c += 16) tmem_st16(base + c, r);
  tmem_st_wait_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  unsigned sink = 0;
  float fs = 0.f;
  const unsigned long long t0 = clock64();
  uint32_t col = (warp >> 2) * 32;      // warpgroups start at different columns
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      tmem_ld16(base + (col & 511), r);
      tmem_ld_wait16(r);
      sink ^= r[0] ^ r[15];
      col += 16;
    } else if (MODE == 1) {
      tmem_ld_32x32b_x32(base + (col & 511), r);
      tmem_ld_wait();
      asm volatile("" : "+r"(r[0]), "+r"(r[31]));
      sink ^= r[0] ^ r[31];
      col += 32;
    } else if (MODE == 2) {
      tmem_ld_16x256b_x2(base + (col & 511), r);
      tmem_ld_16x256b_x2(base + (16u << 16) + (col & 511), r + 8);
      tmem_ld_wait16(r);
      sink ^= r[0] ^ r[15];
      col += 16;
    } else if (MODE == 3) {
      tmem_st16(base + (col & 511), r);
      col += 16;
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float v = __uint_as_float(r[i]);
        v = ex2(v);
        r[i] = __float_as_uint(v * -0.5f);
      }
    } else if (MODE == 5) {
      tmem_ld16(base + (col & 511), r);
      tmem_ld16(base + ((col + 16) & 511), r + 16);
      tmem_ld_wait();
      asm volatile("" : "+r"(r[0]), "+r"(r[31]));
      sink ^= r[0] ^ r[31];
      col += 32;
    }
  }
  if (MODE == 3) tmem_st_wait_all();
  if (MODE == 4) {
#pragma unroll
    for (int i = 0; i < 16; ++i) fs += __uint_as_float(r[i]);
    sink ^= __float_as_uint(fs);
  }
  tc_fence_before();
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x].cycles = t1 - t0; }
  if (sink == 0x12345678u) out[blockIdx.x].sink = sink;
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(s_tmem, 512); }
}

// One elected lane issues `reps` groups of `per_group` UMMAs (M=128, N, K=16, f16) into alternating accumulators, one commit
// per group, and waits for the last commit.  a_tmem: A operand from TMEM (columns 496..503) instead of shared memory.
// small_every > 0: after every group, additionally issue `small_every` UMMAs with N=16 into columns 480..495 (the P.V pattern).
// commits: tcgen05.commit instructions per group (onto scratch barriers nobody waits for; > 1 shows what a commit costs the pipe)
__global__ void __launch_bounds__(128, 1) umma_kernel(int N, int per_group, int reps, int a_tmem, int small_every, Res* out,
                                                      int commits = 0, int commit_split = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ __align__(8) uint64_t s_scratch[4];
  const int warp = threadIdx.x >> 5;
  for (int e = threadIdx.x * 16; e < 64 * 1024; e += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + e) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1); mbar_init(smem_u32(&s_bar[1]), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s_scratch[i]), 1 << 20);     // never completes: arrivals only
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = s_tmem;
  {   // finite A operand in TMEM
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    tmem_st8(tbase + 496 + (((uint32_t)(warp * 32)) << 16), z);
    tmem_st_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc16 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    // A: 128 rows x K=16: core matrices of 8 rows x 16 B, LBO = 128*16 B (second K granule), SBO = 128 B
    const uint64_t adesc = desc_hi(128) | ((uint64_t)((2048u >> 4) & 0x3FFF) << 16) | (uint64_t)(smem_u32(smem) >> 4);
    // B: N rows x K=16 at 16 KB: same canonical layout
    const uint64_t bdesc = desc_hi(128) | ((uint64_t)((4096u >> 4) & 0x3FFF) << 16) | (uint64_t)(smem_u32(smem + 16384) >> 4);
    t0 = clock64();
    for (int g = 0; g < reps; ++g) {
      const uint32_t d = tbase + (g & 1) * 240;
      if (elect_one()) {
        for (int i = 0; i < per_group; ++i) {
          if (a_tmem) umma_f16_ts(d, tbase + 496, bdesc, idesc, i ? 1u : 0u);
          else umma_f16(d, adesc, bdesc, idesc, i ? 1u : 0u);
          if (commit_split && i == per_group / 2) umma_commit(smem_u32(&s_scratch[3]));   // a commit in the middle of a tile
        }
        for (int i = 0; i < commits; ++i) umma_commit(smem_u32(&s_scratch[i & 3]));
        for (int i = 0; i < small_every; ++i) umma_f16_ts(tbase + 480, tbase + 496, bdesc, idesc16, 1u);
        if (g == reps - 1) umma_commit(smem_u32(&s_bar[0]));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&s_bar[0]), 0, 99);
    tc_fence_after();
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x].cycles = t1 - t0;
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

static unsigned long long run_max(Res* d_out, int n) {
  std::vector<Res> h(n);
  CK(cudaMemcpy(h.data(), d_out, n * sizeof(Res), cudaMemcpyDeviceToHost));
  unsigned long long m = 0;
  for (auto& r : h) m = std::max(m, r.cycles);
  return m;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Res* d_out;
  CK(cudaMalloc(&d_out, sms * sizeof(Res)));
  const int iters = 4096;
  printf("# tmem_bench: %d SMs, one CTA per SM, %d iterations per warp\n", sms, iters);
  printf("# kind warps shape  cycles  bytes_per_clk_per_SM  clk_per_128x1_column\n");
  const char* names[] = {"ld 32x32b.x16", "ld 32x32b.x32", "ld 16x256b.x2(x2 halves)", "st 32x32b.x16", "ex2 x16", "ld 32x32b.x16 2-deep"};
  for (int mode = 0; mode < 6; ++mode) {
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {     // first launch warms up
        switch (mode) {
          case 0: ldst_kernel<0><<<sms, warps * 32>>>(iters, d_out); break;
          case 1: ldst_kernel<1><<<sms, warps * 32>>>(iters, d_out); break;
          case 2: ldst_kernel<2><<<sms, warps * 32>>>(iters, d_out); break;
          case 3: ldst_kernel<3><<<sms, warps * 32>>>(iters, d_out); break;
          case 4: ldst_kernel<4><<<sms, warps * 32>>>(iters, d_out); break;
          case 5: ldst_kernel<5><<<sms, warps * 32>>>(iters, d_out); break;
        }
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
      }
      const unsigned long long cyc = run_max(d_out, sms);
      const int cols = (mode == 1 || mode == 5) ? 32 : 16;
      if (mode == 4) {
        const double ex = (double)warps * 32 * 16 * iters;
        printf("mufu  %2d  %-26s %10llu  ex2_per_clk_per_SM=%.2f\n", warps, names[mode], cyc, ex / cyc);
      } else {
        const double bytes = (double)warps * iters * 32.0 * cols * 4.0;
        printf("tmem  %2d  %-26s %10llu  %.1f  %.3f\n", warps, names[mode], cyc, bytes / cyc, 512.0 / (bytes / cyc));
      }
    }
  }
  printf("# umma: N per_group a_tmem small_every -> cycles per group (M=128, K=16, f16), reps=64\n");
  CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const int reps = 64;
  struct Cfg { int N, per, at, small; };
  const Cfg cfgs[] = {{240, 10, 1, 0}, {240, 10, 0, 0}, {240, 10, 1, 15}, {240, 5, 1, 15}, {240, 22, 1, 13}, {240, 22, 1, 0},
                      {16, 16, 1, 0}, {16, 64, 1, 0}, {32, 16, 1, 0}, {64, 16, 1, 0}, {128, 16, 1, 0}, {16, 16, 0, 0},
                      {208, 11, 1, 0}, {208, 11, 1, 13}};
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      umma_kernel<<<sms, 128, 64 * 1024>>>(c.N, c.per, reps, c.at, c.small, d_out);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
    }
    const unsigned long long cyc = run_max(d_out, sms);
    printf("umma  N=%3d per_group=%2d a_tmem=%d small=%2d  total=%llu  per_group=%.1f  per_main_umma=%.1f  (floor 128*N/256 = %.0f)\n",
           c.N, c.per, c.at, c.small, cyc, (double)cyc / reps, (double)cyc / reps / c.per, 128.0 * c.N / 256.0);
  }
  printf("# umma + commits: N=240, 10 per group, A from TMEM; commits per group / one more in the middle of the group\n");
  for (int commits : {0, 1, 2, 3}) for (int split : {0, 1}) {
    for (int rep = 0; rep < 2; ++rep) {
      umma_kernel<<<sms, 128, 64 * 1024>>>(240, 10, reps, 1, 0, d_out, commits, split);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
    }
    const unsigned long long cyc = run_max(d_out, sms);
    printf("commit  commits=%d mid=%d  total=%llu  per_group=%.1f\n", commits, split, cyc, (double)cyc / reps);
  }
  CK(cudaFree(d_out));
  return 0;
}
