// Shared helpers for the cdscore kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CDS_OK 0
#define CDS_ERR_ARG -1
#define CDS_ERR_CUDA -2
#define CDS_ERR_UNSUPPORTED -3

void cds_set_error(const char* fmt, ...);

#define CDS_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      cds_set_error(__VA_ARGS__);                \
      return CDS_ERR_ARG;                        \
    }                                            \
  } while (0)

#define CDS_CHECK_LAUNCH(what)                                               \
  do {                                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      cds_set_error("%s: %s", what, cudaGetErrorString(e__));                \
      return CDS_ERR_CUDA;                                                   \
    }                                                                        \
  } while (0)

#define CDS_LOG2E 1.4426950408889634f

// running (max, sum-exp, weighted sum) in base-2 logits
template <int C>
struct Softmax2 {
  float m, l, acc[C];
  __device__ __forceinline__ void init() {
    m = -INFINITY;
    l = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
  }
  // t = logit * log2(e)
  __device__ __forceinline__ void push(float t, const float* v) {
    if (t > m) {
      float s = exp2f(m - t);  // m = -inf -> 0
      l *= s;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] *= s;
      m = t;
    }
    float e = exp2f(t - m);
    l += e;
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = fmaf(e, v[c], acc[c]);
  }
};
