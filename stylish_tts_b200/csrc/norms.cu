// Normalisation-family kernels: InstanceNorm statistics -> AdaIN affine,
// LayerNorm over the channel axis of (B,C,T), the ConvNeXt front
// (depthwise k7 + LayerNorm + adaptive affine), depthwise Conv1d, GRN scale,
// and the packed style FC.  All fp32; HBM-bound streaming kernels with the time
// axis mapped to consecutive threads (coalesced 128 B per warp access).
#include "common.cuh"

namespace sty {

// ------------------------------------------------------------------ instnorm
// One CTA per (b,c) row.  Two passes over the row (second pass hits L2):
// mean, then sum (x-mean)^2 — same two-pass variance as ATen's CPU batch-norm
// statistics, robust when mean^2 >> var (F0 in Hz).
__global__ void __launch_bounds__(512)
instnorm_affine_kernel(const float* __restrict__ x, int64_t x_bs, int64_t x_cs,
                       const float* __restrict__ gb, int64_t gb_bs, float* __restrict__ scale,
                       float* __restrict__ shift, int C, int T, float eps) {
  __shared__ float red[32];
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const float* __restrict__ row = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
  const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
  const int T4 = vec ? (T >> 2) : 0;
  float s = 0.f;
  for (int i = threadIdx.x; i < T4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(row)[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  for (int i = T4 * 4 + threadIdx.x; i < T; i += blockDim.x) s += row[i];
  const float mean = block_sum(s, red) / (float)T;
  float q = 0.f;
  for (int i = threadIdx.x; i < T4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(row)[i];
    const float a = v.x - mean, bb = v.y - mean, cc = v.z - mean, d = v.w - mean;
    q += (a * a + bb * bb) + (cc * cc + d * d);
  }
  for (int i = T4 * 4 + threadIdx.x; i < T; i += blockDim.x) {
    const float a = row[i] - mean;
    q += a * a;
  }
  const float var = block_sum(q, red) / (float)T;
  if (threadIdx.x == 0) {
    const float rstd = 1.0f / sqrtf(var + eps);
    const float g = gb[(int64_t)b * gb_bs + c];
    const float be = gb[(int64_t)b * gb_bs + C + c];
    const float sc = (1.0f + g) * rstd;
    scale[b * C + c] = sc;
    shift[b * C + c] = be - mean * sc;
  }
}

// ------------------------------------------------------------ chan layernorm
// thread = one (b,t) column; CREG > 0 keeps the column in registers.
template <int CREG>
__global__ void __launch_bounds__(128)
chan_layernorm_kernel(const float* __restrict__ xin, const float* __restrict__ resin, int64_t x_bs,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      int64_t g_bs, int g_plus_one, float* __restrict__ yout, int64_t y_bs,
                      const float* __restrict__ mask, int C, int T, float eps, int act, int64_t x_cs,
                      int64_t y_cs) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* __restrict__ x = xin + (int64_t)b * x_bs;
  const float* __restrict__ res = resin ? resin + (int64_t)b * x_bs : nullptr;
  float* __restrict__ y = yout + (int64_t)b * y_bs;
  const int64_t base = t;
  const float* __restrict__ g = gamma + (int64_t)b * g_bs;
  const float* __restrict__ be = beta + (int64_t)b * g_bs;
  const float m = mask ? mask[(int64_t)b * T + t] : 1.f;
  const float invC = 1.0f / (float)C;
  if constexpr (CREG > 0) {
    float v[CREG];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CREG; ++c) {
      v[c] = x[base + (int64_t)c * x_cs];
      if (res) v[c] += res[base + (int64_t)c * x_cs];
      s += v[c];
    }
    const float mean = s * invC;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CREG; ++c) {
      const float d = v[c] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(q * invC + eps);
#pragma unroll
    for (int c = 0; c < CREG; ++c) {
      const float gg = g_plus_one ? 1.0f + g[c] : g[c];
      float o = (v[c] - mean) * rstd * gg + be[c];
      y[base + (int64_t)c * y_cs] = act_apply(o, act) * m;
    }
  } else {
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      float v = x[base + (int64_t)c * x_cs];
      if (res) v += res[base + (int64_t)c * x_cs];
      s += v;
    }
    const float mean = s * invC;
    float q = 0.f;
    for (int c = 0; c < C; ++c) {
      float v = x[base + (int64_t)c * x_cs];
      if (res) v += res[base + (int64_t)c * x_cs];
      const float d = v - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(q * invC + eps);
    for (int c = 0; c < C; ++c) {
      float v = x[base + (int64_t)c * x_cs];
      if (res) v += res[base + (int64_t)c * x_cs];
      const float gg = g_plus_one ? 1.0f + g[c] : g[c];
      float o = (v - mean) * rstd * gg + be[c];
      y[base + (int64_t)c * y_cs] = act_apply(o, act) * m;
    }
  }
}

// ---------------------------------------------------------------- dwconv + LN
// thread = one (b,t) column.  d[c] = bias[c] + sum_k w[c,k] x[b,c,t+k-3]
template <int CREG>
__global__ void __launch_bounds__(128)
dwconv_ln_kernel(const float* __restrict__ x, int64_t x_bs, const float* __restrict__ w,
                 const float* __restrict__ bias, const float* __restrict__ gb, int64_t gb_bs,
                 float* __restrict__ yout, int64_t y_bs, int C, int T, float eps) {
  constexpr int K = 7, P = 3;
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* __restrict__ xb = x + (int64_t)b * x_bs;
  float* __restrict__ y = yout + (int64_t)b * y_bs;
  const float* __restrict__ g = gb + (int64_t)b * gb_bs;
  const float invC = 1.0f / (float)C;
  auto dw = [&](int c) {
    const float* __restrict__ xr = xb + (int64_t)c * T;
    float a = bias[c];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int u = t + k - P;
      if (u >= 0 && u < T) a = fmaf(w[c * K + k], xr[u], a);
    }
    return a;
  };
  if constexpr (CREG > 0) {
    float d[CREG];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CREG; ++c) {
      d[c] = dw(c);
      s += d[c];
    }
    const float mean = s * invC;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < CREG; ++c) {
      const float e = d[c] - mean;
      q = fmaf(e, e, q);
    }
    const float rstd = 1.0f / sqrtf(q * invC + eps);
#pragma unroll
    for (int c = 0; c < CREG; ++c)
      y[(int64_t)c * T + t] = (1.0f + g[c]) * ((d[c] - mean) * rstd) + g[C + c];
  } else {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += dw(c);
    const float mean = s * invC;
    float q = 0.f;
    for (int c = 0; c < C; ++c) {
      const float e = dw(c) - mean;
      q = fmaf(e, e, q);
    }
    const float rstd = 1.0f / sqrtf(q * invC + eps);
    for (int c = 0; c < C; ++c)
      y[(int64_t)c * T + t] = (1.0f + g[c]) * ((dw(c) - mean) * rstd) + g[C + c];
  }
}

// --------------------------------------------------------------- column norm
// LayerNorm over channels of a (C x TT) tile, optionally preceded by the depthwise k7
// conv of the ConvNeXt block.  The tile (+ halo) is staged once in shared memory with
// coalesced loads; thread (tl, cs) owns time step tl and the channel slice {cs, cs+CS, ..}
// (values stay in registers), partial sums meet in shared memory.  Used for C <= 256
// where one-thread-per-column has too little parallelism (text encoder T=258, F-rate).
template <bool DW, int TT>
__global__ void __launch_bounds__(256)
colnorm_kernel(const float* __restrict__ xin, const float* __restrict__ resin, int64_t x_bs,
               const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ gamma, const float* __restrict__ beta, int64_t g_bs,
               int g_plus_one, float* __restrict__ yout, int64_t y_bs, const float* __restrict__ mask,
               int C, int T, float eps, int act) {
  constexpr int P = DW ? 3 : 0, K = 7;
  constexpr int W = TT + 2 * P + 1;  // +1: odd row pitch
  constexpr int CS = 256 / TT;       // channel slices
  constexpr int NC = 32;             // channels per thread: C <= 32 * CS (256 for TT=32, 64 for TT=128)
  extern __shared__ float sm[];
  float* xs = sm;                    // [C][W]
  float* red = sm + (size_t)C * W;   // [CS][TT]
  const int b = blockIdx.y, t0 = blockIdx.x * TT, tid = threadIdx.x;
  const float* __restrict__ x = xin + (int64_t)b * x_bs;
  const float* __restrict__ res = resin ? resin + (int64_t)b * x_bs : nullptr;
  for (int idx = tid; idx < C * (TT + 2 * P); idx += 256) {
    const int c = idx / (TT + 2 * P), j = idx - c * (TT + 2 * P);
    const int t = t0 - P + j;
    float v = 0.f;
    if (t >= 0 && t < T) {
      v = x[(int64_t)c * T + t];
      if (res) v += res[(int64_t)c * T + t];
    }
    xs[c * W + j] = v;
  }
  __syncthreads();
  const int tl = tid % TT, cs = tid / TT;
  const int t = t0 + tl;
  float d[NC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = cs + i * CS;
    float a = 0.f;
    if (c < C) {
      if (DW) {
        a = bias[c];
#pragma unroll
        for (int k = 0; k < K; ++k) a = fmaf(w[c * K + k], xs[c * W + tl + k], a);
      } else {
        a = xs[c * W + tl];
      }
    }
    d[i] = a;
    s += a;
  }
  red[cs * TT + tl] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < CS; ++i) tot += red[i * TT + tl];
  const float mean = tot / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int c = cs + i * CS;
    const float e = (c < C) ? d[i] - mean : 0.f;
    q = fmaf(e, e, q);
  }
  __syncthreads();
  red[cs * TT + tl] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int i = 0; i < CS; ++i) qt += red[i * TT + tl];
  const float rstd = 1.0f / sqrtf(qt / (float)C + eps);
  if (t < T) {
    const float* __restrict__ g = gamma + (int64_t)b * g_bs;
    const float* __restrict__ be = beta + (int64_t)b * g_bs;
    const float m = mask ? mask[(int64_t)b * T + t] : 1.f;
    float* __restrict__ y = yout + (int64_t)b * y_bs + t;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int c = cs + i * CS;
      if (c < C) {
        const float gg = g_plus_one ? 1.0f + g[c] : g[c];
        const float o = (d[i] - mean) * rstd * gg + be[c];
        y[(int64_t)c * T] = act_apply(o, act) * m;
      }
    }
  }
}

// -------------------------------------------------------------------- dwconv
__global__ void __launch_bounds__(256)
dwconv1d_kernel(const float* __restrict__ x, int64_t x_bs, int64_t x_cs,
                const float* __restrict__ w, const float* __restrict__ bias,
                const float* __restrict__ post_scale, const float* __restrict__ post_shift,
                float* __restrict__ y, int64_t y_bs, int64_t y_cs, int C, int T, int K,
                int pad_left, int act) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float* __restrict__ xr = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
  float a = bias ? bias[c] : 0.f;
  for (int k = 0; k < K; ++k) {
    const int u = t + k - pad_left;
    if (u >= 0 && u < T) a = fmaf(w[c * K + k], xr[u], a);
  }
  if (post_scale) a = fmaf(a, post_scale[c], post_shift ? post_shift[c] : 0.f);
  y[(int64_t)b * y_bs + (int64_t)c * y_cs + t] = act_apply(a, act);
}

// ----------------------------------------------------------------- GRN scale
__global__ void __launch_bounds__(256)
grn_scale_kernel(const float* __restrict__ sumsq, const float* __restrict__ gamma,
                 float* __restrict__ scale, int J) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int j = threadIdx.x; j < J; j += blockDim.x) s += sqrtf(sumsq[(int64_t)b * J + j]);
  const float mean = block_sum(s, red) / (float)J;
  const float inv = 1.0f / (mean + 1e-6f);
  for (int j = threadIdx.x; j < J; j += blockDim.x)
    scale[(int64_t)b * J + j] = 1.0f + gamma[j] * (sqrtf(sumsq[(int64_t)b * J + j]) * inv);
}

// --------------------------------------------------------------- linear rows
// one warp per (b, j) output: lanes stride the I inputs.
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ s, const float* __restrict__ W,
                   const float* __restrict__ bias, float* __restrict__ out, int B, int I, int J) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (int64_t)B * J) return;
  const int b = (int)(warp / J), j = (int)(warp % J);
  float a = 0.f;
  for (int i = lane; i < I; i += 32) a = fmaf(W[(int64_t)j * I + i], s[(int64_t)b * I + i], a);
  a = warp_sum(a);
  if (lane == 0) out[(int64_t)b * J + j] = a + (bias ? bias[j] : 0.f);
}

}  // namespace sty

using namespace sty;

extern "C" int sty_instnorm_affine_fwd(const float* x, int64_t x_bs, int64_t x_cs, const float* gb,
                                       int64_t gb_bs, float* scale, float* shift, int B, int C,
                                       int T, float eps, sty_stream_t stream) {
  STY_REQUIRE(x && gb && scale && shift, "instnorm_affine: null pointer");
  STY_REQUIRE(B > 0 && C > 0 && T > 0, "instnorm_affine: bad shape");
  const int threads = T >= 4096 ? 512 : (T >= 512 ? 256 : 64);
  instnorm_affine_kernel<<<B * C, threads, 0, as_stream(stream)>>>(x, x_bs, x_cs, gb, gb_bs, scale,
                                                                    shift, C, T, eps);
  STY_CHECK_LAUNCH("instnorm_affine");
  return STY_OK;
}

extern "C" int sty_chan_layernorm_pitched_fwd(const float* x, const float* res, int64_t x_bs, int64_t x_cs,
                                              const float* gamma, const float* beta, int64_t g_bs,
                                              int g_plus_one, float* y, int64_t y_bs, int64_t y_cs,
                                              const float* mask, int B, int C, int T, float eps, int act,
                                              sty_stream_t stream) {
  STY_REQUIRE(x && gamma && beta && y, "chan_layernorm: null pointer");
  STY_REQUIRE(B > 0 && C > 0 && T > 0 && x_cs >= T && y_cs >= T, "chan_layernorm: bad shape");
  dim3 grid(cdiv(T, 128), B);
  cudaStream_t st = as_stream(stream);
#define LAUNCH(CR)                                                                               \
  chan_layernorm_kernel<CR><<<grid, 128, 0, st>>>(x, res, x_bs, gamma, beta, g_bs, g_plus_one, y, \
                                                  y_bs, mask, C, T, eps, act, x_cs, y_cs)
  if (C == 32) LAUNCH(32);
  else if (C == 64) LAUNCH(64);
  else if (C <= 256 && x_cs == T && y_cs == T) {
    constexpr int TT = 32;
    const size_t smem = ((size_t)C * (TT + 1) + 256) * sizeof(float);
    dim3 g2(cdiv(T, TT), B);
    colnorm_kernel<false, TT><<<g2, 256, smem, st>>>(x, res, x_bs, nullptr, nullptr, gamma, beta, g_bs,
                                                    g_plus_one, y, y_bs, mask, C, T, eps, act);
  } else LAUNCH(0);
#undef LAUNCH
  STY_CHECK_LAUNCH("chan_layernorm");
  return STY_OK;
}

extern "C" int sty_chan_layernorm_fwd(const float* x, const float* res, int64_t x_bs,
                                      const float* gamma, const float* beta, int64_t g_bs,
                                      int g_plus_one, float* y, int64_t y_bs, const float* mask,
                                      int B, int C, int T, float eps, int act,
                                      sty_stream_t stream) {
  return sty_chan_layernorm_pitched_fwd(x, res, x_bs, T, gamma, beta, g_bs, g_plus_one, y, y_bs, T, mask, B, C, T,
                                        eps, act, stream);
}

extern "C" int sty_dwconv_ln_fwd(const float* x, int64_t x_bs, const float* w, const float* bias,
                                 const float* gb, int64_t gb_bs, float* y, int64_t y_bs, int B,
                                 int C, int T, float eps, sty_stream_t stream) {
  STY_REQUIRE(x && w && bias && gb && y, "dwconv_ln: null pointer");
  STY_REQUIRE(B > 0 && C > 0 && T > 0, "dwconv_ln: bad shape");
  dim3 grid(cdiv(T, 128), B);
  cudaStream_t st = as_stream(stream);
  if (C <= 64) {
    // smem-staged tile, 128 time steps x (2 channel slices): one global read per element
    constexpr int TT = 128;
    const size_t smem = ((size_t)C * (TT + 7) + 256) * sizeof(float);
    dim3 g2(cdiv(T, TT), B);
    colnorm_kernel<true, TT><<<g2, 256, smem, st>>>(x, nullptr, x_bs, w, bias, gb, gb + C, gb_bs, 1, y, y_bs,
                                                   nullptr, C, T, eps, STY_ACT_NONE);
  } else if (C <= 256) {
    constexpr int TT = 32;
    const size_t smem = ((size_t)C * (TT + 7) + 256) * sizeof(float);
    dim3 g2(cdiv(T, TT), B);
    colnorm_kernel<true, TT><<<g2, 256, smem, st>>>(x, nullptr, x_bs, w, bias, gb, gb + C, gb_bs, 1, y, y_bs,
                                                   nullptr, C, T, eps, STY_ACT_NONE);
  } else {
    dwconv_ln_kernel<0><<<grid, 128, 0, st>>>(x, x_bs, w, bias, gb, gb_bs, y, y_bs, C, T, eps);
  }
  STY_CHECK_LAUNCH("dwconv_ln");
  return STY_OK;
}

extern "C" int sty_dwconv1d_fwd(const float* x, int64_t x_bs, int64_t x_cs, const float* w,
                                const float* bias, const float* post_scale,
                                const float* post_shift, float* y, int64_t y_bs, int64_t y_cs,
                                int B, int C, int T, int K, int pad_left, int act,
                                sty_stream_t stream) {
  STY_REQUIRE(x && w && y, "dwconv1d: null pointer");
  STY_REQUIRE(B > 0 && C > 0 && T > 0 && K > 0 && pad_left >= 0 && pad_left < K, "dwconv1d: bad shape");
  STY_REQUIRE(C <= 65535 && B <= 65535, "dwconv1d: grid too large");
  dim3 grid(cdiv(T, 256), C, B);
  dwconv1d_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_bs, x_cs, w, bias, post_scale,
                                                       post_shift, y, y_bs, y_cs, C, T, K, pad_left,
                                                       act);
  STY_CHECK_LAUNCH("dwconv1d");
  return STY_OK;
}

extern "C" int sty_grn_scale_fwd(const float* sumsq, const float* gamma, float* scale, int B, int J,
                                 sty_stream_t stream) {
  STY_REQUIRE(sumsq && gamma && scale && B > 0 && J > 0, "grn_scale: bad argument");
  grn_scale_kernel<<<B, 256, 0, as_stream(stream)>>>(sumsq, gamma, scale, J);
  STY_CHECK_LAUNCH("grn_scale");
  return STY_OK;
}

extern "C" int sty_linear_rows_fwd(const float* s, const float* W, const float* bias, float* out,
                                   int B, int I, int J, sty_stream_t stream) {
  STY_REQUIRE(s && W && out && B > 0 && I > 0 && J > 0, "linear_rows: bad argument");
  const int64_t warps = (int64_t)B * J;
  linear_rows_kernel<<<cdiv(warps * 32, 256), 256, 0, as_stream(stream)>>>(s, W, bias, out, B, I, J);
  STY_CHECK_LAUNCH("linear_rows");
  return STY_OK;
}

// ------------------------------------------------------------ AdaIN affine from accumulated moments
namespace sty {
__global__ void moments_affine_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq,
                                      const float* __restrict__ gb, int64_t gb_bs, float* __restrict__ scale,
                                      float* __restrict__ shift, int B, int C, float invT, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i - b * C;
  const float mean = sum[i] * invT;
  const float var = fmaxf(fmaf(-mean, mean, sumsq[i] * invT), 0.f);
  const float sc = (1.0f + gb[(int64_t)b * gb_bs + c]) / sqrtf(var + eps);
  scale[i] = sc;
  shift[i] = gb[(int64_t)b * gb_bs + C + c] - mean * sc;
}
}  // namespace sty

extern "C" int sty_moments_affine_fwd(const float* sum, const float* sumsq, const float* gb, int64_t gb_bs,
                                      float* scale, float* shift, int B, int C, int T, float eps,
                                      sty_stream_t stream) {
  using namespace sty;
  STY_REQUIRE(sum && sumsq && gb && scale && shift && B > 0 && C > 0 && T > 0, "moments_affine: bad argument");
  moments_affine_kernel<<<cdiv((int64_t)B * C, 128), 128, 0, as_stream(stream)>>>(sum, sumsq, gb, gb_bs, scale, shift,
                                                                               B, C, 1.0f / (float)T, eps);
  STY_CHECK_LAUNCH("moments_affine");
  return STY_OK;
}
