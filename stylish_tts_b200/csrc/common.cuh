// Shared device/host helpers for libstylish_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stylish_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libstylish_b200 targets sm_100a only"
#endif

namespace sty {

void set_error(const char* fmt, ...);

#define STY_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      sty::set_error(__VA_ARGS__);             \
      return STY_ERR_BAD_ARG;                  \
    }                                          \
  } while (0)

#define STY_CHECK_LAUNCH(name)                                                  \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      sty::set_error("%s: CUDA error: %s%s", name, cudaGetErrorString(e__),     \
                     e__ == cudaErrorMemoryAllocation ? " (out of memory)" : ""); \
      return STY_ERR_CUDA;                                                      \
    }                                                                           \
  } while (0)

static inline cudaStream_t as_stream(sty_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- device math
__device__ __forceinline__ float act_apply(float v, int act, float alpha) {
  switch (act) {
    case STY_ACT_RELU:
      return fmaxf(v, 0.f);
    case STY_ACT_LEAKY02:
      return v > 0.f ? v : 0.2f * v;
    case STY_ACT_SNAKE: {
      float s = sinf(alpha * v);
      return v + (1.0f / alpha) * (s * s);
    }
    case STY_ACT_SWISH:
      return v / (1.0f + expf(-v));
    default:
      return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` is >= 32 floats of shared memory; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) red[0] = r;
  }
  __syncthreads();
  return red[0];
}

}  // namespace sty
