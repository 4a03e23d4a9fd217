// Shared helpers for the ess_b200 CUDA sources.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ess_b200.h"

void essb_set_error(const char* fmt, ...);

#define ESSB_REQUIRE(cond, ...)              \
  do {                                       \
    if (!(cond)) {                           \
      essb_set_error(__VA_ARGS__);           \
      return ESSB_ERR_ARG;                   \
    }                                        \
  } while (0)

#define ESSB_LAUNCH_CHECK(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      essb_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return ESSB_ERR_LAUNCH;                                                     \
    }                                                                             \
  } while (0)

static inline bool essb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float essb_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
