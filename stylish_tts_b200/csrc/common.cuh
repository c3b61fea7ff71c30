// Shared device/host helpers for libstylish_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
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
// sin^2(y), the Snake activation's inner function; it sits in conv prologues / epilogues where it was 40 % of
// all instructions of the fused ConvNeXt kernel (ncu source view, profiles/r01_ncu_pw1_source_blocks.txt).
// Default: one MUFU.SIN (sin.approx, |abs error| ~ 1e-6 after the 1/2pi range scaling) and a multiply —
// 3 instructions.  STY_PRECISE_SNAKE=1 selects the Cody-Waite + degree-11 polynomial (|error| < 1e-7, ~20
// instructions) that the first parity runs used.
#ifndef STY_PRECISE_SNAKE
#define STY_PRECISE_SNAKE 0
#endif
__device__ __forceinline__ float sin_sq(float y) {
#if STY_PRECISE_SNAKE
  const float k = rintf(y * 0.318309886183790672f);
  float r = fmaf(k, -3.14159274101257324f, y);
  r = fmaf(k, 8.74227800037247e-08f, r);
  const float r2 = r * r;
  float p = fmaf(r2, -2.50521083854417e-08f, 2.75573192239859e-06f);
  p = fmaf(p, r2, -1.98412698412698e-04f);
  p = fmaf(p, r2, 8.33333333333333e-03f);
  p = fmaf(p, r2, -1.66666666666667e-01f);
  const float s = fmaf(r * r2, p, r);
  return s * s;
#else
  const float s = __sinf(y);
  return s * s;
#endif
}

// `alpha` / `inv_alpha` are only read for STY_ACT_SNAKE.
__device__ __forceinline__ float act_apply(float v, int act, float alpha, float inv_alpha) {
  switch (act) {
    case STY_ACT_RELU:
      return fmaxf(v, 0.f);
    case STY_ACT_LEAKY02:
      return v > 0.f ? v : 0.2f * v;
    case STY_ACT_SNAKE:
      return fmaf(inv_alpha, sin_sq(alpha * v), v);
    case STY_ACT_SWISH:
      return v / (1.0f + expf(-v));
    case STY_ACT_GELU:
      return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    default:
      return v;
  }
}
__device__ __forceinline__ float act_apply(float v, int act) { return act_apply(v, act, 1.f, 1.f); }

// ---------------------------------------------------------------- dropout masks
// Stateless counter-based keep mask: keep(seed, site, idx) = hash >= p * 2^24, the same three inputs in the
// forward and the backward kernel, so no mask is ever stored.  `seed` is read from DEVICE memory at run time:
// a captured CUDA graph draws new masks on every replay.  The integer hash is restated in numpy by
// oracle/dropout_oracle.py (bit-exact), which is how the dropout sites are parity-tested.
struct DropSpec {
  const unsigned long long* seed;  // nullptr: dropout off
  uint32_t site, thresh;
  float inv_keep;
};
static inline DropSpec make_drop(const sty_dropout* d) {
  DropSpec s{nullptr, 0u, 0u, 1.f};
  if (d && d->seed && d->p > 0.f) {
    s.seed = reinterpret_cast<const unsigned long long*>(d->seed);
    s.site = d->site;
    s.thresh = (uint32_t)llrint((double)d->p * 16777216.0);
    s.inv_keep = 1.0f / (1.0f - d->p);
  }
  return s;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool drop_keep(unsigned long long seed, uint32_t site, unsigned long long idx,
                                          uint32_t thresh) {
  uint32_t x = mix32((uint32_t)idx ^ (uint32_t)seed);
  x = mix32(x ^ ((uint32_t)(idx >> 32) + site * 0x9E3779B9u + (uint32_t)(seed >> 32)));
  return (x >> 8) >= thresh;
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

// ---------------------------------------------------------------- vectorised row streaming
// Walk N equally long rows (and optionally an output row) of T floats with the block's threads:
// out[t] = f(t, {rows[0][t], ..., rows[N-1][t]}).  When all pointers share their 16-byte misalignment the bulk
// goes through 128-bit loads / stores (<= 3 scalar elements at either end), otherwise everything is scalar.
// Streaming kernels with one 4-byte load per thread and iteration keep ~32 KB in flight per SM, about half of
// what HBM3e needs (tools/bench_bwd_stream.py: 1.1-2.3 TB/s before).
template <int N, bool WRITE, typename F>
__device__ __forceinline__ void rows_apply(const float* const (&rows)[N], float* out, int T, int tid, int nth,
                                           F&& f) {
  const unsigned mis = (unsigned)((uintptr_t)rows[0] & 15u);
  bool alike = true;
#pragma unroll
  for (int k = 1; k < N; ++k) alike &= ((unsigned)((uintptr_t)rows[k] & 15u) == mis);
  if (WRITE) alike &= ((unsigned)((uintptr_t)out & 15u) == mis);
  int head = T;
  if (alike) {
    head = (int)(((16u - mis) & 15u) >> 2);
    if (head > T) head = T;
  }
  const int nv = (T - head) >> 2;
  const int tail = head + 4 * nv;
  for (int t = tid; t < head; t += nth) {
    float v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = rows[k][t];
    const float o = f(t, v);
    if (WRITE) out[t] = o;
  }
#pragma unroll 2
  for (int i = tid; i < nv; i += nth) {
    float4 q[N];
#pragma unroll
    for (int k = 0; k < N; ++k) q[k] = reinterpret_cast<const float4*>(rows[k] + head)[i];
    const int t = head + 4 * i;
    float v[N];
    float4 o;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = q[k].x;
    o.x = f(t, v);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = q[k].y;
    o.y = f(t + 1, v);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = q[k].z;
    o.z = f(t + 2, v);
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = q[k].w;
    o.w = f(t + 3, v);
    if (WRITE) reinterpret_cast<float4*>(out + head)[i] = o;
  }
  for (int t = tail + tid; t < T; t += nth) {
    float v[N];
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = rows[k][t];
    const float o = f(t, v);
    if (WRITE) out[t] = o;
  }
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
