// Device-side pieces of the adversarial losses: LeakyReLU + space-to-depth (strided
// convolutions of the spectrogram discriminators become stride-1 tensor-core convolutions on the rearranged
// input), the LSGAN sums, and TPRLS (median of dr - dg by radix select, masked squared deviations) without any
// host synchronisation.  Reference: train/models/discriminator.py:13-69, train/losses.py:166-373.
#include "common.cuh"

namespace sty {
namespace {

// y[n, c*s + p, u] = leaky(x[n, c, s*u + p])  (0 past the end); s = 1: plain LeakyReLU
__global__ void leaky_s2d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int C, int W,
                                     int W2, int s, float slope) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % W2);
    const int64_t r = i / W2;       // n * C*s + c*s + p
    const int p = (int)(r % s);
    const int64_t nc = r / s;       // n * C + c
    const int w = u * s + p;
    float v = 0.f;
    if (w < W) {
      v = x[nc * W + w];
      v = v > 0.f ? v : slope * v;
    }
    y[i] = v;
  }
}

// dx[n, c, w] = dy[n, c*s + w % s, w / s] * (x > 0 ? 1 : slope)
__global__ void leaky_s2d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx,
                                     int64_t total, int C, int W, int W2, int s, float slope) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int64_t nc = i / W;
    const float g = dy[(nc * s + (w % s)) * W2 + w / s];
    dx[i] = x[i] > 0.f ? g : slope * g;
  }
}

// y[r, t] = (x ? x[r, t] : 1) * g[r] * mul  — the squeeze-excite style gate of ContextFreeDiscriminator
// (discriminator.py:167-168: x * attn) and the broadcast of a per-row gradient
__global__ void row_scale_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y,
                                 int64_t total, int T, float mul) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = g[i / T] * mul;
    y[i] = x ? x[i] * s : s;
  }
}

// y[r] = mul * sum_{j < P} x[r*P + j]: pooled value of short segments (AdaptiveAvgPool1d over one window of the
// gapped waveform-discriminator layout); one thread per segment, P % 4 == 0
__global__ void segment_sum_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int P, float mul) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float4* p = reinterpret_cast<const float4*>(x + r * P);
    float acc = 0.f;
    for (int j = 0; j < P / 4; ++j) {
      const float4 v = p[j];
      acc += (v.x + v.y) + (v.z + v.w);
    }
    y[r] = acc * mul;
  }
}

// ---- sums of squares: out[0] += sum (c - x)^2
__global__ void sqdiff_sum_kernel(const float* __restrict__ x, int64_t n, float c, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = c - x[i];
    acc = fmaf(d, d, acc);
  }
  for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// ---- radix select of the lower median of d = a - b (torch.median semantics: sorted[(n-1)/2])
__device__ __forceinline__ uint32_t ordered_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// state: [0] prefix key bits found so far, [1] remaining rank inside the prefix bucket, hist[2048] after it
template <int SHIFT, int BITS>
__global__ void select_hist_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                   const uint32_t* __restrict__ state, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[1 << BITS];
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) h[i] = 0;
  __syncthreads();
  const uint32_t prefix = state[0];
  constexpr uint32_t hi_mask = (SHIFT + BITS >= 32) ? 0u : (0xffffffffu << (SHIFT + BITS));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = ordered_key(a[i] - b[i]);
    if ((k & hi_mask) == (prefix & hi_mask)) atomicAdd(&h[(k >> SHIFT) & ((1u << BITS) - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}
template <int SHIFT, int BITS>
__global__ void select_pick_kernel(uint32_t* __restrict__ state, uint32_t* __restrict__ hist) {
  // one block: find the bucket containing the remaining rank, update prefix / rank, clear the histogram
  __shared__ uint32_t cum[1 << BITS];
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) cum[i] = hist[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t rank = state[1], run = 0;
    int bucket = (1 << BITS) - 1;
    for (int i = 0; i < (1 << BITS); ++i) {
      if (run + cum[i] > rank) {
        bucket = i;
        break;
      }
      run += cum[i];
    }
    state[0] |= ((uint32_t)bucket << SHIFT);
    state[1] = rank - run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) hist[i] = 0;
}
__global__ void select_init_kernel(uint32_t* state, uint32_t rank) {
  state[0] = 0;
  state[1] = rank;
}

// sums[0] = sum_{a < b + m} (a - b - m)^2, sums[1] = count, sums[2] = sum_{mask} (a - b - m);  m = key_to_float(state[0])
__global__ void tprls_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                  const uint32_t* __restrict__ state, float* __restrict__ sums, float* __restrict__ med) {
  const float m = key_to_float(state[0]);
  if (blockIdx.x == 0 && threadIdx.x == 0) *med = m;
  float s2 = 0.f, cnt = 0.f, s1 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float av = a[i], bv = b[i];
    if (av < bv + m) {
      const float d = (av - bv) - m;
      s2 = fmaf(d, d, s2);
      s1 += d;
      cnt += 1.f;
    }
  }
  for (int off = 16; off; off >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[0], s2);
    atomicAdd(&sums[1], cnt);
    atomicAdd(&sums[2], s1);
  }
}

// d(l_rel)/d(a - b): coef[0] * mask * (a - b - m) + (first element equal to the median) * coef[1]; da = +, db = -
__global__ void tprls_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                 const float* __restrict__ med, const float* __restrict__ coef, float* __restrict__ da,
                                 float* __restrict__ db, int* __restrict__ flag) {
  const float m = *med, c0 = coef[0], c1 = coef[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float av = a[i], bv = b[i];
    const float d = av - bv;
    float g = (av < bv + m) ? c0 * (d - m) : 0.f;
    if (d == m && atomicExch(flag, 1) == 0) g += c1;  // the median element (one of them on ties)
    if (da) da[i] = g;
    if (db) db[i] = -g;
  }
}

int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace
}  // namespace sty

using namespace sty;

extern "C" int sty_leaky_s2d_fwd(const float* x, float* y, int64_t N, int C, int W, int s, float slope,
                                 sty_stream_t stream) {
  STY_REQUIRE(x && y && N > 0 && C > 0 && W > 0 && s >= 1, "leaky_s2d: bad argument");
  const int W2 = (W + s - 1) / s;
  const int64_t total = N * C * s * W2;
  leaky_s2d_fwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(x, y, total, C, W, W2, s, slope);
  STY_CHECK_LAUNCH("leaky_s2d_fwd");
  return STY_OK;
}

extern "C" int sty_leaky_s2d_bwd(const float* dy, const float* x, float* dx, int64_t N, int C, int W, int s,
                                 float slope, sty_stream_t stream) {
  STY_REQUIRE(dy && x && dx && N > 0 && C > 0 && W > 0 && s >= 1, "leaky_s2d_bwd: bad argument");
  const int W2 = (W + s - 1) / s;
  const int64_t total = N * C * W;
  leaky_s2d_bwd_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(dy, x, dx, total, C, W, W2, s, slope);
  STY_CHECK_LAUNCH("leaky_s2d_bwd");
  return STY_OK;
}

extern "C" int sty_row_scale_fwd(const float* x, const float* g, float* y, int64_t rows, int T, float mul,
                                 sty_stream_t stream) {
  STY_REQUIRE(g && y && rows > 0 && T > 0, "row_scale: bad argument");
  const int64_t total = rows * T;
  row_scale_kernel<<<blocks_for(total), 256, 0, as_stream(stream)>>>(x, g, y, total, T, mul);
  STY_CHECK_LAUNCH("row_scale");
  return STY_OK;
}

extern "C" int sty_segment_sum_fwd(const float* x, float* y, int64_t rows, int P, float mul, sty_stream_t stream) {
  STY_REQUIRE(x && y && rows > 0 && P > 0 && P % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "segment_sum: bad argument (P %% 4 == 0, 16-byte aligned rows)");
  segment_sum_kernel<<<blocks_for(rows), 256, 0, as_stream(stream)>>>(x, y, rows, P, mul);
  STY_CHECK_LAUNCH("segment_sum");
  return STY_OK;
}

extern "C" int sty_sqdiff_sum_fwd(const float* x, int64_t n, float c, float* out, sty_stream_t stream) {
  STY_REQUIRE(x && out && n > 0, "sqdiff_sum: bad argument");
  sqdiff_sum_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(x, n, c, out);
  STY_CHECK_LAUNCH("sqdiff_sum");
  return STY_OK;
}

extern "C" int64_t sty_tprls_workspace_bytes(void) { return (2 + 2048) * 4 + 16; }

extern "C" int sty_tprls_fwd(const float* a, const float* b, int64_t n, void* workspace, float* sums, float* median,
                             sty_stream_t stream) {
  STY_REQUIRE(a && b && workspace && sums && median && n > 0 && n < (1ll << 32), "tprls: bad argument");
  cudaStream_t st = as_stream(stream);
  uint32_t* state = reinterpret_cast<uint32_t*>(workspace);
  uint32_t* hist = state + 2;
  if (cudaMemsetAsync(workspace, 0, (size_t)sty_tprls_workspace_bytes(), st) != cudaSuccess ||
      cudaMemsetAsync(sums, 0, 3 * sizeof(float), st) != cudaSuccess) {
    set_error("tprls: memset failed");
    return STY_ERR_CUDA;
  }
  select_init_kernel<<<1, 1, 0, st>>>(state, (uint32_t)((n - 1) / 2));
  const int g = blocks_for(n);
  select_hist_kernel<21, 11><<<g, 256, 0, st>>>(a, b, n, state, hist);
  select_pick_kernel<21, 11><<<1, 256, 0, st>>>(state, hist);
  select_hist_kernel<10, 11><<<g, 256, 0, st>>>(a, b, n, state, hist);
  select_pick_kernel<10, 11><<<1, 256, 0, st>>>(state, hist);
  select_hist_kernel<0, 10><<<g, 256, 0, st>>>(a, b, n, state, hist);
  select_pick_kernel<0, 10><<<1, 256, 0, st>>>(state, hist);
  tprls_sums_kernel<<<g, 256, 0, st>>>(a, b, n, state, sums, median);
  STY_CHECK_LAUNCH("tprls_fwd");
  return STY_OK;
}

extern "C" int sty_tprls_bwd(const float* a, const float* b, int64_t n, const float* median, const float* coef,
                             float* da, float* db, int* flag, sty_stream_t stream) {
  STY_REQUIRE(a && b && median && coef && flag && (da || db) && n > 0, "tprls_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(flag, 0, sizeof(int), st) != cudaSuccess) {
    set_error("tprls_bwd: memset failed");
    return STY_ERR_CUDA;
  }
  tprls_bwd_kernel<<<blocks_for(n), 256, 0, st>>>(a, b, n, median, coef, da, db, flag);
  STY_CHECK_LAUNCH("tprls_bwd");
  return STY_OK;
}
