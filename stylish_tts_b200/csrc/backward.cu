// Backward (training) kernels of the hot path: Conv1d weight gradient, the
// backward of the fused conv prologue (mask / AdaIN-BatchNorm affine / activation),
// GRN + Snake, channel LayerNorm, depthwise Conv1d, GLU, embedding, the spectral
// head + conv-iSTFT, the packed style FC, and the fused AdamW update.
// Data gradients of Conv1d reuse sty_conv1d_fwd with transposed, tap-reversed weights
// (a stride-1 'same' convolution is its own adjoint up to that relabelling).
// See include/stylish_b200.h for the semantics of every entry point.
#include <math.h>

#include "common.cuh"

namespace sty {
namespace {

// derivative of the activation at pre-activation a
__device__ __forceinline__ float act_grad(float a, int act, float alpha) {
  switch (act) {
    case STY_ACT_RELU:
      return a > 0.f ? 1.f : 0.f;
    case STY_ACT_LEAKY02:
      return a > 0.f ? 1.f : 0.2f;
    case STY_ACT_SNAKE:  // d/da (a + sin^2(alpha a)/alpha) = 1 + sin(2 alpha a); MUFU.SIN like the forward
      return 1.f + __sinf(2.f * alpha * a);
    case STY_ACT_SWISH: {
      const float s = 1.f / (1.f + expf(-a));
      return s * (1.f + a * (1.f - s));
    }
    case STY_ACT_GELU: {
      const float cdf = 0.5f * (1.f + erff(a * 0.70710678118654752440f));
      return cdf + a * 0.3989422804014327f * expf(-0.5f * a * a);
    }
    default:
      return 1.f;
  }
}

// sin^2(theta) and sin(2 theta) from one MUFU.SIN + MUFU.COS pair: everything the Snake backward needs
struct SnakeTerms {
  float s2, sd;
};
__device__ __forceinline__ SnakeTerms snake_terms(float theta) {
  float sn, cs;
  __sincosf(theta, &sn, &cs);
  return {sn * sn, 2.f * sn * cs};
}
// d/d alpha of snake(a; alpha) = a sin(2 alpha a)/alpha - sin^2(alpha a)/alpha^2
__device__ __forceinline__ float snake_dalpha(float a, float inv_alpha, const SnakeTerms& st) {
  return (a * st.sd - st.s2 * inv_alpha) * inv_alpha;
}

__device__ __forceinline__ float prologue_value(float xv, float m, float sc, float sh, int act, float al) {
  if (m < 0.f) return 0.f;  // negative mask value: zero AFTER the prologue (gapped layouts)
  float w = fmaf(xv * m, sc, sh);
  if (act == STY_ACT_SNAKE) return fmaf(1.f / al, sin_sq(al * w), w);
  return act_apply(w, act);
}

// ------------------------------------------------------------------ conv1d weight gradient
// CTA: 32*CI_R input channels x 8*CO_R output channels x all KT taps, grid-strided over
// (batch, 128-step time chunks); partial sums are added to dw with atomics at the end.
// lane -> ci, warp -> CO_R consecutive co.  The output-gradient tile is staged time-major
// so one broadcast 128-bit shared load feeds 4 output channels; the input tile (prologue
// applied while staging, same as the forward) has an odd pitch so the lane-strided reads
// are conflict free.
template <int KT, int CO_R>
__global__ void __launch_bounds__(256)
conv1d_wgrad_kernel(const sty_conv1d_wgrad_args p, const int n_tchunks, const int up) {
  constexpr int TT = 128, CO_T = 8 * CO_R, GP = CO_T + 4;
  extern __shared__ __align__(16) float smem[];
  float* gs = smem;            // [TT][GP]
  float* us = smem + TT * GP;  // [32][up]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_ci_tiles = (p.CI + 31) / 32;
  const int co0 = (blockIdx.y / n_ci_tiles) * CO_T;
  const int ci0 = (blockIdx.y % n_ci_tiles) * 32;
  const int UT = TT + (KT - 1) * p.dil;
  float acc[CO_R][KT];
#pragma unroll
  for (int i = 0; i < CO_R; ++i)
#pragma unroll
    for (int k = 0; k < KT; ++k) acc[i][k] = 0.f;

  const int64_t total = (int64_t)p.B * n_tchunks;
  for (int64_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    const int b = (int)(chunk / n_tchunks);
    const int t0 = (int)(chunk - (int64_t)b * n_tchunks) * TT;
    // ---- output gradient tile: g = out_scale * mask_o * dy
    const float* __restrict__ dyb = p.dy + (int64_t)b * p.dy_bs;
    const float* __restrict__ om = p.out_mask ? p.out_mask + (int64_t)b * p.T : nullptr;
    for (int idx = tid; idx < CO_T * TT; idx += 256) {
      const int c = idx / TT, tt = idx - c * TT;
      const int co = co0 + c, t = t0 + tt;
      float v = 0.f;
      if (co < p.CO && t < p.T) {
        v = dyb[(int64_t)co * p.dy_cs + t] * p.out_scale;
        if (om) v *= om[t];
      }
      gs[tt * GP + c] = v;
    }
    // ---- input tile with the forward prologue
    const float* __restrict__ xb = p.x + (int64_t)b * p.x_bs;
    const float* __restrict__ im = p.in_mask ? p.in_mask + (int64_t)b * p.T : nullptr;
    for (int idx = tid; idx < 32 * UT; idx += 256) {
      const int c = idx / UT, tt = idx - c * UT;
      const int ci = ci0 + c, t = t0 - p.pad + tt;
      float v = 0.f;
      if (ci < p.CI && t >= 0 && t < p.T) {
        const float sc = p.in_scale ? p.in_scale[(int64_t)b * p.CI + ci] : 1.f;
        const float sh = p.in_shift ? p.in_shift[(int64_t)b * p.CI + ci] : 0.f;
        const float al = p.in_alpha ? p.in_alpha[ci] : 1.f;
        v = prologue_value(xb[(int64_t)ci * p.x_cs + t], im ? im[t] : 1.f, sc, sh, p.in_act, al);
      }
      us[c * up + tt] = v;
    }
    __syncthreads();
    const float* __restrict__ urow = us + lane * up;
    const float* __restrict__ grow = gs + warp * CO_R;
#pragma unroll 2
    for (int tt = 0; tt < TT; ++tt) {
      float g[CO_R];
#pragma unroll
      for (int i = 0; i < CO_R; i += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(grow + tt * GP + i);
        g[i] = g4.x; g[i + 1] = g4.y; g[i + 2] = g4.z; g[i + 3] = g4.w;
      }
#pragma unroll
      for (int k = 0; k < KT; ++k) {
        const float u = urow[tt + k * p.dil];
#pragma unroll
        for (int i = 0; i < CO_R; ++i) acc[i][k] = fmaf(g[i], u, acc[i][k]);
      }
    }
    __syncthreads();
  }
  const int ci = ci0 + lane;
  if (ci < p.CI) {
#pragma unroll
    for (int i = 0; i < CO_R; ++i) {
      const int co = co0 + warp * CO_R + i;
      if (co < p.CO) {
        float* __restrict__ d = p.dw + ((int64_t)co * p.CI + ci) * KT;
#pragma unroll
        for (int k = 0; k < KT; ++k) atomicAdd(d + k, acc[i][k]);
      }
    }
  }
}

template <int KT, int CO_R>
int launch_wgrad(const sty_conv1d_wgrad_args& a, cudaStream_t st) {
  constexpr int TT = 128, CO_T = 8 * CO_R, GP = CO_T + 4;
  int up = TT + (KT - 1) * a.dil;
  if ((up & 1) == 0) ++up;
  const size_t smem = ((size_t)TT * GP + 32 * (size_t)up) * sizeof(float);
  STY_REQUIRE(smem <= 200 * 1024, "conv1d_wgrad: footprint too large (K=%d dil=%d)", a.K, a.dil);
  auto kern = conv1d_wgrad_kernel<KT, CO_R>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int n_tchunks = cdiv(a.T, TT);
  const int tiles = cdiv(a.CO, CO_T) * cdiv(a.CI, 32);
  int sms = sty_device_sm_count();
  if (sms <= 0) sms = 148;
  int64_t splits = ((int64_t)4 * sms + tiles - 1) / tiles;  // ~4 CTAs per SM in total
  const int64_t total = (int64_t)a.B * n_tchunks;
  if (splits > total) splits = total;
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)splits, (unsigned)tiles);
  kern<<<grid, 256, smem, st>>>(a, n_tchunks, up);
  STY_CHECK_LAUNCH("conv1d_wgrad");
  return STY_OK;
}

// ------------------------------------------------------------------ per-channel sums
// out[c] += scale * sum_{b,t} mask[b,t] * x[b,c,t]
__global__ void __launch_bounds__(256)
channel_sum_kernel(const float* __restrict__ x, int64_t x_bs, int64_t x_cs, const float* __restrict__ mask,
                   float* __restrict__ out, int B, int T, float scale) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  float s = 0.f;
  // grid.y <= 64 blocks per channel walk the batch (the image convs of the style encoder have B ~ 2600 rows of
  // T ~ 800: one block and one atomic per (b, c) row was 210 K blocks)
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const float* row = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
    const float* m = mask ? mask + (int64_t)b * T : nullptr;
    if (m) {
      const float* const rows[2] = {row, m};
      rows_apply<2, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[2]) {
        s = fmaf(v[0], v[1], s);
        return 0.f;
      });
    } else {
      const float* const rows[1] = {row};
      rows_apply<1, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[1]) {
        s += v[0];
        return 0.f;
      });
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out + c, s * scale);
}

// out[b,c] = sum_t a[b,c,t] * bb[b,c,t]
__global__ void __launch_bounds__(256)
row_dot_kernel(const float* __restrict__ a, const float* __restrict__ bb, float* __restrict__ out, int T) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const float* __restrict__ ar = a + r * T;
  const float* __restrict__ br = bb + r * T;
  float s = 0.f;
  const float* const rows[2] = {ar, br};
  rows_apply<2, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[2]) {
    s = fmaf(v[0], v[1], s);
    return 0.f;
  });
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[r] = s;
}

// ------------------------------------------------------------------ row moments
// mean[b,c], var[b,c] (biased, two-pass) over T
__global__ void __launch_bounds__(512)
row_moments_kernel(const float* __restrict__ x, int64_t x_bs, int64_t x_cs, float* __restrict__ mean,
                   float* __restrict__ var, int C, int T) {
  __shared__ float red[32];
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const float* __restrict__ row = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
  float s = 0.f;
  const float* const rows[1] = {row};
  rows_apply<1, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[1]) {
    s += v[0];
    return 0.f;
  });
  const float mu = block_sum(s, red) / (float)T;
  float q = 0.f;
  rows_apply<1, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[1]) {
    const float a = v[0] - mu;
    q = fmaf(a, a, q);
    return 0.f;
  });
  const float v = block_sum(q, red) / (float)T;
  if (threadIdx.x == 0) {
    mean[blockIdx.x] = mu;
    var[blockIdx.x] = v;
  }
}

// ------------------------------------------------------------------ prologue backward
// a = scale*(x*m)+shift, u = act(a); g_a = dxp * act'(a)
// reduce: sums[b,c,0] = sum g_a, [1] = sum g_a*(x*m - center), [2] = sum dxp * d snake/d alpha
__global__ void __launch_bounds__(512)
prologue_bwd_reduce_kernel(const float* __restrict__ dxp, const float* __restrict__ x, int64_t x_bs, int64_t x_cs,
                           const float* __restrict__ scale, const float* __restrict__ shift,
                           const float* __restrict__ alpha, const float* __restrict__ mask,
                           const float* __restrict__ center, float* __restrict__ sums, int C, int T, int act) {
  __shared__ float red[32];
  const int b = blockIdx.x / C, c = blockIdx.x % C;
  const float* __restrict__ xr = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
  const float* __restrict__ gr = dxp + (int64_t)blockIdx.x * T;
  const float* __restrict__ m = mask ? mask + (int64_t)b * T : nullptr;
  const float sc = scale ? scale[blockIdx.x] : 1.f, sh = shift ? shift[blockIdx.x] : 0.f;
  const float al = alpha ? alpha[c] : 1.f;
  const float ce = center ? center[blockIdx.x] : 0.f;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  const float inv_al = 1.f / al;
  auto body = [&](float xm, float g) {
    const float a = fmaf(xm, sc, sh);
    float ga;
    if (act == STY_ACT_SNAKE) {
      const SnakeTerms st = snake_terms(al * a);
      ga = g * (1.f + st.sd);
      s2 = fmaf(g, snake_dalpha(a, inv_al, st), s2);
    } else {
      ga = g * act_grad(a, act, al);
    }
    s0 += ga;
    s1 = fmaf(ga, xm - ce, s1);
  };
  if (m) {
    const float* const rows[3] = {xr, gr, m};
    rows_apply<3, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[3]) {
      body(v[0] * v[2], v[1]);
      return 0.f;
    });
  } else {
    const float* const rows[2] = {xr, gr};
    rows_apply<2, false>(rows, nullptr, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[2]) {
      body(v[0], v[1]);
      return 0.f;
    });
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    sums[(int64_t)blockIdx.x * 3] = s0;
    sums[(int64_t)blockIdx.x * 3 + 1] = s1;
    sums[(int64_t)blockIdx.x * 3 + 2] = s2;
  }
}

// apply: dx = g_a*scale*m + c0[b,c] + c1[b,c]*x (+ add); one block = a 1024-element span of one (b,c) row
constexpr int kApplySpan = 1024;
__global__ void __launch_bounds__(256)
prologue_bwd_apply_kernel(const float* __restrict__ dxp, const float* __restrict__ x, int64_t x_bs, int64_t x_cs,
                          const float* __restrict__ scale, const float* __restrict__ shift,
                          const float* __restrict__ alpha, const float* __restrict__ mask,
                          const float* __restrict__ c0, const float* __restrict__ c1,
                          const float* __restrict__ add, int64_t add_bs, int64_t add_cs, float* __restrict__ dx,
                          int64_t dx_bs, int64_t dx_cs, int C, int T, int act) {
  const int bc = blockIdx.x;
  const int b = bc / C, c = bc % C;
  const int t0 = blockIdx.y * kApplySpan;
  const int n = min(kApplySpan, T - t0);
  const float sc = scale ? scale[bc] : 1.f, sh = shift ? shift[bc] : 0.f;
  const float al = alpha ? alpha[c] : 1.f;
  const float k0 = c0 ? c0[bc] : 0.f, k1 = c1 ? c1[bc] : 0.f;
  const float* xr = x + (int64_t)b * x_bs + (int64_t)c * x_cs + t0;
  const float* gr = dxp + (int64_t)bc * T + t0;
  const float* mr = mask ? mask + (int64_t)b * T + t0 : nullptr;
  const float* ar = add ? add + (int64_t)b * add_bs + (int64_t)c * add_cs + t0 : nullptr;
  float* out = dx + (int64_t)b * dx_bs + (int64_t)c * dx_cs + t0;
  auto body = [&](float xv, float g, float m, float addv) {
    const float a = fmaf(xv * m, sc, sh);
    return fmaf(k1, xv, g * act_grad(a, act, al) * sc * m + k0) + addv;
  };
  if (!mr && !ar) {  // the S-rate case: AdaIN / Snake prologues of the generator blocks
    const float* const rows[2] = {xr, gr};
    rows_apply<2, true>(rows, out, n, threadIdx.x, blockDim.x,
                        [&](int, const float (&v)[2]) { return body(v[0], v[1], 1.f, 0.f); });
  } else if (mr && !ar) {
    const float* const rows[3] = {xr, gr, mr};
    rows_apply<3, true>(rows, out, n, threadIdx.x, blockDim.x,
                        [&](int, const float (&v)[3]) { return body(v[0], v[1], v[2], 0.f); });
  } else if (!mr) {
    const float* const rows[3] = {xr, gr, ar};
    rows_apply<3, true>(rows, out, n, threadIdx.x, blockDim.x,
                        [&](int, const float (&v)[3]) { return body(v[0], v[1], 1.f, v[2]); });
  } else {
    const float* const rows[4] = {xr, gr, mr, ar};
    rows_apply<4, true>(rows, out, n, threadIdx.x, blockDim.x,
                        [&](int, const float (&v)[4]) { return body(v[0], v[1], v[2], v[3]); });
  }
}

// ------------------------------------------------------------------ GRN + Snake backward (apply)
// hb = snake(h); d_hb = g_u*gs[b,j] + kc[b,j]*hb; d_h = d_hb * snake'(h); dalpha[j] += sum d_hb * dsnake/dalpha
// one CTA per (b,j) row; d_h may alias g_u.
__global__ void __launch_bounds__(512)
grn_snake_bwd_kernel(const float* g_u, const float* __restrict__ h, const float* __restrict__ gs,
                     const float* __restrict__ kc, const float* __restrict__ alpha, float* d_h,
                     float* __restrict__ dalpha, int J, int T) {
  __shared__ float red[32];
  const int j = blockIdx.x % J;
  const int64_t off = (int64_t)blockIdx.x * T;
  const float s = gs[blockIdx.x], k = kc[blockIdx.x], al = alpha[j], inv = 1.f / al;
  float da = 0.f;
  const float* const rows[2] = {h + off, g_u + off};
  rows_apply<2, true>(rows, d_h + off, T, threadIdx.x, blockDim.x, [&](int, const float (&v)[2]) {
    const float hv = v[0];
    const SnakeTerms st = snake_terms(al * hv);
    const float hb = fmaf(inv, st.s2, hv);
    const float dhb = fmaf(v[1], s, k * hb);
    da = fmaf(dhb, snake_dalpha(hv, inv, st), da);
    return dhb * (1.f + st.sd);
  });
  da = block_sum(da, red);
  if (threadIdx.x == 0) atomicAdd(dalpha + j, da);
}

// ------------------------------------------------------------------ channel LayerNorm backward
// thread = one (b,t) column.  v = x (+res); n = (v-mean)*rstd; z = G*n + be; y = act(z)*mask
// dz = dy*mask*act'(z); dn = dz*G; dv = rstd*(dn - mean(dn) - n*mean(dn*n))
// dgb[b*dg_bs + c] += sum_t dz*n ; dgb[b*dg_bs + C + c] += sum_t dz   (dg_bs = 0: shared)
__global__ void __launch_bounds__(128)
chan_layernorm_bwd_kernel(const float* __restrict__ xin, const float* __restrict__ resin, int64_t x_bs,
                          const float* __restrict__ gamma, const float* __restrict__ beta, int64_t g_bs,
                          int g_plus_one, const float* __restrict__ dyin, const float* __restrict__ mask,
                          float* __restrict__ dvout, float* __restrict__ dgb, int64_t dg_bs, int C, int T,
                          float eps, int act) {
  extern __shared__ float sacc[];  // [2*C] sums, then per warp two [32][33] reduction tiles
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const bool ok = t < T;
  const float* __restrict__ x = xin + (int64_t)b * x_bs + (ok ? t : 0);
  const float* __restrict__ res = resin ? resin + (int64_t)b * x_bs + (ok ? t : 0) : nullptr;
  const float* __restrict__ dy = dyin + (int64_t)b * C * T + (ok ? t : 0);
  float* __restrict__ dv = dvout + (int64_t)b * C * T + (ok ? t : 0);
  const float* __restrict__ g = gamma + (int64_t)b * g_bs;
  const float* __restrict__ be = beta + (int64_t)b * g_bs;
  const float m = (mask && ok) ? mask[(int64_t)b * T + t] : 1.f;
  const float invC = 1.f / (float)C;
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    float v = x[(int64_t)c * T];
    if (res) v += res[(int64_t)c * T];
    s += v;
  }
  const float mean = s * invC;
  float q = 0.f;
  for (int c = 0; c < C; ++c) {
    float v = x[(int64_t)c * T];
    if (res) v += res[(int64_t)c * T];
    v -= mean;
    q = fmaf(v, v, q);
  }
  const float rstd = 1.f / sqrtf(q * invC + eps);
  float m1 = 0.f, m2 = 0.f;
  // d(gamma), d(beta) need sums over t, i.e. over the lanes: per 32 channels every lane parks its two values in
  // a padded [channel][lane] tile of shared memory and then sums ONE channel over the 32 lanes — 4 shared-memory
  // operations per value instead of two 5-step shuffle reductions (20 instructions)
  float* ta = sacc + 2 * C + (threadIdx.x >> 5) * (2 * 32 * 33);
  float* tb = ta + 32 * 33;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int nc = min(32, C - c0);
    for (int i = 0; i < nc; ++i) {
      const int c = c0 + i;
      float v = x[(int64_t)c * T];
      if (res) v += res[(int64_t)c * T];
      const float n = (v - mean) * rstd;
      const float G = g_plus_one ? 1.f + g[c] : g[c];
      float dz = ok ? dy[(int64_t)c * T] * m : 0.f;
      if (act != STY_ACT_NONE) dz *= act_grad(fmaf(G, n, be[c]), act, 1.f);
      const float dn = dz * G;
      m1 += dn;
      m2 = fmaf(dn, n, m2);
      ta[i * 33 + lane] = dz * n;
      tb[i * 33 + lane] = dz;
    }
    __syncwarp();
    if (lane < nc) {
      float a = 0.f, bsum = 0.f;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) {
        a += ta[lane * 33 + l];
        bsum += tb[lane * 33 + l];
      }
      atomicAdd(&sacc[c0 + lane], a);
      atomicAdd(&sacc[C + c0 + lane], bsum);
    }
    __syncwarp();
  }
  m1 *= invC;
  m2 *= invC;
  if (ok) {
    for (int c = 0; c < C; ++c) {
      float v = x[(int64_t)c * T];
      if (res) v += res[(int64_t)c * T];
      const float n = (v - mean) * rstd;
      const float G = g_plus_one ? 1.f + g[c] : g[c];
      float dz = dy[(int64_t)c * T] * m;
      if (act != STY_ACT_NONE) dz *= act_grad(fmaf(G, n, be[c]), act, 1.f);
      dv[(int64_t)c * T] = rstd * (dz * G - m1 - n * m2);
    }
  }
  __syncthreads();
  float* __restrict__ d = dgb + (int64_t)b * dg_bs;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(d + i, sacc[i]);
}

// ------------------------------------------------------------------ depthwise conv backward
// dx[b,c,t] = sum_k w[c,k]*dy[b,c,t-k+pad] (+add); dw[c,k] += sum dy[t]*x[t+k-pad]; db[c] += sum dy
template <int KMAX>
__global__ void __launch_bounds__(256)
dwconv1d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t x_bs, int64_t x_cs,
                    const float* __restrict__ w, float* __restrict__ dx, int64_t dx_bs, int64_t dx_cs,
                    float* __restrict__ dw, float* __restrict__ db, int C, int T, int K, int pad) {
  __shared__ float sacc[KMAX + 1];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* __restrict__ dr = dy + ((int64_t)b * C + c) * T;
  const float* __restrict__ xr = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
  float* __restrict__ dxr = dx ? dx + (int64_t)b * dx_bs + (int64_t)c * dx_cs : nullptr;
  for (int i = threadIdx.x; i <= KMAX; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float wk[KMAX], aw[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    wk[k] = k < K ? w[c * K + k] : 0.f;
    aw[k] = 0.f;
  }
  float ab = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float g = dr[t];
    ab += g;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const int u = t - k + pad;  // dy index feeding dx[t] through tap k
        if (u >= 0 && u < T) v = fmaf(wk[k], dr[u], v);
        const int xi = t + k - pad;
        if (xi >= 0 && xi < T) aw[k] = fmaf(g, xr[xi], aw[k]);
      }
    }
    if (dxr) dxr[t] = v;
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if (k < K) {
      const float a = warp_sum(aw[k]);
      if (lane == 0) atomicAdd(&sacc[k], a);
    }
  }
  ab = warp_sum(ab);
  if (lane == 0) atomicAdd(&sacc[KMAX], ab);
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(dw + c * K + k, sacc[k]);
  if (threadIdx.x == 0 && db) atomicAdd(db + c, sacc[KMAX]);
}

// Windowed variant for the S-rate ConvNeXt depthwise convs (K = 7 exactly): a thread owns R = 4 consecutive
// time steps and loads the dy / x windows they share once (20 loads per 4 outputs instead of 56).
template <int K, int PAD, int R>
__global__ void __launch_bounds__(256)
dwconv1d_bwd_win_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t x_bs, int64_t x_cs,
                        const float* __restrict__ w, float* __restrict__ dx, int64_t dx_bs, int64_t dx_cs,
                        float* __restrict__ dw, float* __restrict__ db, int C, int T) {
  static_assert(PAD >= 0 && PAD <= K - 1, "window layout");
  constexpr int pad = PAD;
  __shared__ float sacc[K + 1];
  const int c = blockIdx.x, b = blockIdx.y;
  const float* __restrict__ dr = dy + ((int64_t)b * C + c) * T;
  const float* __restrict__ xr = x + (int64_t)b * x_bs + (int64_t)c * x_cs;
  float* __restrict__ dxr = dx ? dx + (int64_t)b * dx_bs + (int64_t)c * dx_cs : nullptr;
  for (int i = threadIdx.x; i <= K; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  float wk[K], aw[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    wk[k] = w[c * K + k];
    aw[k] = 0.f;
  }
  float ab = 0.f;
  constexpr int W = R + K - 1;
  for (int t0 = threadIdx.x * R; t0 < T; t0 += blockDim.x * R) {
    float wd[W], wx[W];
    const int bd = t0 - (K - 1) + pad, bx = t0 - pad;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int ud = bd + j, ux = bx + j;
      wd[j] = (ud >= 0 && ud < T) ? dr[ud] : 0.f;
      wx[j] = (ux >= 0 && ux < T) ? xr[ux] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float g = wd[r + (K - 1) - PAD];  // dy[t0 + r] (0 beyond T)
      ab += g;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        v = fmaf(wk[k], wd[r - k + (K - 1)], v);  // dy[t - k + pad]
        aw[k] = fmaf(g, wx[r + k], aw[k]);        // x[t + k - pad]
      }
      if (dxr && t0 + r < T) dxr[t0 + r] = v;
    }
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float a = warp_sum(aw[k]);
    if (lane == 0) atomicAdd(&sacc[k], a);
  }
  ab = warp_sum(ab);
  if (lane == 0) atomicAdd(&sacc[K], ab);
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) atomicAdd(dw + c * K + k, sacc[k]);
  if (threadIdx.x == 0 && db) atomicAdd(db + c, sacc[K]);
}

// ------------------------------------------------------------------ small elementwise backward
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int C, int T) {
  const int64_t n = (int64_t)C * T;
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = x[(int64_t)b * 2 * n + i], g = x[(int64_t)b * 2 * n + n + i];
  const float s = 1.f / (1.f + expf(-g));
  const float d = dy[(int64_t)b * n + i];
  dx[(int64_t)b * 2 * n + i] = d * s;
  dx[(int64_t)b * 2 * n + n + i] = d * a * s * (1.f - s);
}

// y (B, C, T*s) pixel-shuffled -> x (B, C*s, T): x[b, c*s + r, t] = y[b, c, t*s + r]
__global__ void __launch_bounds__(256)
unshuffle_kernel(const float* __restrict__ y, float* __restrict__ x, int CO, int T, int s) {
  const int b = blockIdx.z, co = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int c = co / s, r = co - c * s;
  x[((int64_t)b * CO + co) * T + t] = y[((int64_t)b * (CO / s) + c) * T * s + (int64_t)t * s + r];
}

// d_emb[tok[b,t], c] += dx[b,c,t] * scale * (t < len[b])
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ lengths,
                 const float* __restrict__ dx, float* __restrict__ d_emb, int T, int C, int n_tokens, float scale) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T || (lengths && t >= lengths[b])) return;
  int64_t tok = tokens[(int64_t)b * T + t];
  tok = tok < 0 ? 0 : (tok >= n_tokens ? n_tokens - 1 : tok);
  atomicAdd(d_emb + tok * C + c, dx[((int64_t)b * C + c) * T + t] * scale);
}

// C[b] (M,N) = A[b] (M,K) @ Bt[b] (N,K)^T
__global__ void __launch_bounds__(256)
bmm_nt_kernel(const float* __restrict__ A, int64_t a_bs, const float* __restrict__ Bt, int64_t b_bs,
              float* __restrict__ Cm, int64_t c_bs, int M, int N, int K) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;
  const float* __restrict__ Ab = A + (int64_t)b * a_bs;
  const float* __restrict__ Bb = Bt + (int64_t)b * b_bs;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int idx = tid; idx < BM * BK; idx += 256) {
      const int kk = idx % BK, mm = idx / BK;
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? Ab[(int64_t)m * K + k] : 0.f;
      const int n = n0 + mm;
      Bs[kk][mm] = (n < N && k < K) ? Bb[(int64_t)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tm + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tn + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* __restrict__ Cb = Cm + (int64_t)b * c_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + 16 * j;
      if (n < N) Cb[(int64_t)m * N + n] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------ packed style FC backward
// dW[j,i] = sum_b dh[b,j]*s[b,i]; dbias[j] = sum_b dh[b,j]
__global__ void __launch_bounds__(256)
linear_rows_bwd_w_kernel(const float* __restrict__ dh, const float* __restrict__ s, float* __restrict__ dW,
                         float* __restrict__ dbias, int B, int I, int J) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)J * I) return;
  const int j = (int)(idx / I), i = (int)(idx - (int64_t)j * I);
  float a = 0.f, bsum = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = dh[(int64_t)b * J + j];
    a = fmaf(d, s[(int64_t)b * I + i], a);
    bsum += d;
  }
  dW[idx] = a;
  if (i == 0) dbias[j] = bsum;
}

// ds[b,i] = sum_j dh[b,j]*W[j,i]     one CTA per b, threads stride over j, I <= 256
__global__ void __launch_bounds__(256)
linear_rows_bwd_s_kernel(const float* __restrict__ dh, const float* __restrict__ W, float* __restrict__ ds, int I,
                         int J) {
  extern __shared__ float part[];  // [8][I]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lane -> i (strided), warp -> j slice
  for (int i0 = 0; i0 < I; i0 += 32) {
    const int i = i0 + lane;
    float a = 0.f;
    if (i < I)
      for (int j = warp; j < J; j += 8) a = fmaf(dh[(int64_t)b * J + j], W[(int64_t)j * I + i], a);
    if (i < I) part[warp * I + i] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += part[w * I + i];
    ds[(int64_t)b * I + i] = a;
  }
}

// ------------------------------------------------------------------ spectral head + iSTFT backward
// thread = one frame f in [0,S]; dwave[n'] = dout*(1-out^2) on the trimmed range.
template <int NFFT, int HOP, int BINS, int FT>
__global__ void __launch_bounds__(FT)
istft_head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                      const float* __restrict__ logamp, int64_t logamp_bs, const float* __restrict__ real,
                      const float* __restrict__ imag, int64_t ri_bs, const float* __restrict__ basis_re,
                      const float* __restrict__ basis_im, float* __restrict__ d_logamp, float* __restrict__ d_real,
                      float* __restrict__ d_imag, int64_t dri_bs, int S) {
  extern __shared__ __align__(16) float sm[];
  float* bre = sm;                  // [BINS][NFFT]
  float* bim = bre + BINS * NFFT;   // [BINS][NFFT]
  float* dw = bim + BINS * NFFT;    // [FT*HOP + NFFT] padded-signal gradient
  const int b = blockIdx.y, tid = threadIdx.x;
  const int f0 = blockIdx.x * FT;
  const int L = S * HOP;
  for (int i = tid; i < BINS * NFFT; i += FT) {
    bre[i] = basis_re[i];
    bim[i] = basis_im[i];
  }
  // padded index n' = HOP*f + n; output sample index = n' - NFFT/2
  for (int i = tid; i < FT * HOP + NFFT; i += FT) {
    const int64_t o = (int64_t)f0 * HOP + i - NFFT / 2;
    float v = 0.f;
    if (o >= 0 && o < L) {
      const float y = out[(int64_t)b * L + o];
      v = dout[(int64_t)b * L + o] * (1.f - y * y);
    }
    dw[i] = v;
  }
  __syncthreads();
  const int f = f0 + tid;
  if (f > S) return;
  float wv[NFFT];
#pragma unroll
  for (int n = 0; n < NFFT; ++n) wv[n] = dw[tid * HOP + n];
  const int fs = f < S ? f : S - 1;  // frame S is the replicate pad of frame S-1
  for (int k = 0; k < BINS; ++k) {
    float dre = 0.f, dim = 0.f;
#pragma unroll
    for (int n = 0; n < NFFT; ++n) {
      dre = fmaf(wv[n], bre[k * NFFT + n], dre);
      dim = fmaf(wv[n], bim[k * NFFT + n], dim);
    }
    dim = -dim;
    const int64_t o = (int64_t)k * S + fs;
    const float mag = expf(logamp[(int64_t)b * logamp_bs + o]);
    const float re = real[(int64_t)b * ri_bs + o], im = imag[(int64_t)b * ri_bs + o];
    const float r2 = re * re + im * im;
    float c = 1.f, s = 0.f, inv_r = 0.f;  // atan2(0,0) = 0
    if (r2 > 0.f) {
      inv_r = rsqrtf(r2);
      c = re * inv_r;
      s = im * inv_r;
    }
    const float dla = (dre * c + dim * s) * mag;
    const float dph = mag * (dim * c - dre * s);
    const float dr = -dph * s * inv_r, di = dph * c * inv_r;
    const int64_t oo = (int64_t)b * BINS * S + o, oi = (int64_t)b * dri_bs + o;
    if (f < S - 1) {
      d_logamp[oo] = dla;
      d_real[oi] = dr;
      d_imag[oi] = di;
    } else {  // frames S-1 and S both land on column S-1
      atomicAdd(d_logamp + oo, dla);
      atomicAdd(d_real + oi, dr);
      atomicAdd(d_imag + oi, di);
    }
  }
}

// ------------------------------------------------------------------ fused AdamW (flat arena)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, float gscale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float pi = p[i];
    pi *= 1.f - lr * wd;
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2 + eps;  // bc2 = sqrt(1 - b2^t)
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// same update with the learning rate and the step count read from device memory (hyper = [lr, step]) so that the
// launch can be replayed from a CUDA graph while both advance
__global__ void __launch_bounds__(256)
adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, const float* __restrict__ hyper, float b1, float b2, float eps, float wd, float gscale) {
  const float lr = hyper[0], step = hyper[1];
  const float bc1 = 1.f - powf(b1, step), bc2 = sqrtf(1.f - powf(b2, step));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float pi = p[i] * (1.f - lr * wd);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - (lr / bc1) * (mi / (sqrtf(vi) / bc2 + eps));
  }
}

}  // namespace
}  // namespace sty

using namespace sty;

namespace sty {
bool conv1d_wgrad_umma_eligible(const sty_conv1d_wgrad_args& a);  // wgrad_umma.cu
int conv1d_wgrad_umma_launch(const sty_conv1d_wgrad_args& a, cudaStream_t st);
}  // namespace sty

extern "C" int sty_conv1d_wgrad(const sty_conv1d_wgrad_args* a, sty_stream_t stream) {
  STY_REQUIRE(a && a->x && a->dy && a->dw, "conv1d_wgrad: null pointer");
  STY_REQUIRE(a->B > 0 && a->CI > 0 && a->CO > 0 && a->T > 0 && a->dil >= 1, "conv1d_wgrad: bad shape");
  STY_REQUIRE(2 * a->pad == (a->K - 1) * a->dil, "conv1d_wgrad: only 'same' padding is supported");
  STY_REQUIRE(a->in_act != STY_ACT_SNAKE || a->in_alpha, "conv1d_wgrad: snake prologue needs in_alpha");
  cudaStream_t st = as_stream(stream);
  if (conv1d_wgrad_umma_eligible(*a)) return conv1d_wgrad_umma_launch(*a, st);
  switch (a->K) {
    case 1: return launch_wgrad<1, 8>(*a, st);
    case 3: return launch_wgrad<3, 8>(*a, st);
    case 5: return launch_wgrad<5, 4>(*a, st);
    case 7: return launch_wgrad<7, 4>(*a, st);
    case 9: return launch_wgrad<9, 4>(*a, st);  // first layer of the spectrogram discriminators (3x9)
    case 11: return launch_wgrad<11, 4>(*a, st);
    case 21: return launch_wgrad<21, 4>(*a, st);
    default:
      set_error("conv1d_wgrad: kernel size %d not built (1,3,5,7,9,11,21)", a->K);
      return STY_ERR_BAD_ARG;
  }
}

extern "C" int sty_channel_sum(const float* x, int64_t x_bs, int64_t x_cs, const float* mask, float* out, int B,
                               int C, int T, float scale, sty_stream_t stream) {
  STY_REQUIRE(x && out && B > 0 && C > 0 && T > 0 && B <= 65535, "channel_sum: bad argument");
  dim3 grid(C, B < 64 ? B : 64);
  channel_sum_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_bs, x_cs, mask, out, B, T, scale);
  STY_CHECK_LAUNCH("channel_sum");
  return STY_OK;
}

extern "C" int sty_row_dot(const float* a, const float* b, float* out, int64_t rows, int T, sty_stream_t stream) {
  STY_REQUIRE(a && b && out && rows > 0 && T > 0, "row_dot: bad argument");
  row_dot_kernel<<<(unsigned)rows, 256, 0, as_stream(stream)>>>(a, b, out, T);
  STY_CHECK_LAUNCH("row_dot");
  return STY_OK;
}

extern "C" int sty_row_moments(const float* x, int64_t x_bs, int64_t x_cs, float* mean, float* var, int B, int C,
                               int T, sty_stream_t stream) {
  STY_REQUIRE(x && mean && var && B > 0 && C > 0 && T > 0, "row_moments: bad argument");
  row_moments_kernel<<<B * C, 512, 0, as_stream(stream)>>>(x, x_bs, x_cs, mean, var, C, T);
  STY_CHECK_LAUNCH("row_moments");
  return STY_OK;
}

extern "C" int sty_prologue_bwd_reduce(const float* dxp, const float* x, int64_t x_bs, int64_t x_cs,
                                       const float* scale, const float* shift, const float* alpha,
                                       const float* mask, const float* center, float* sums, int B, int C, int T,
                                       int act, sty_stream_t stream) {
  STY_REQUIRE(dxp && x && sums && B > 0 && C > 0 && T > 0, "prologue_bwd_reduce: bad argument");
  STY_REQUIRE(act != STY_ACT_SNAKE || alpha, "prologue_bwd_reduce: snake needs alpha");
  prologue_bwd_reduce_kernel<<<B * C, 512, 0, as_stream(stream)>>>(dxp, x, x_bs, x_cs, scale, shift, alpha, mask,
                                                                   center, sums, C, T, act);
  STY_CHECK_LAUNCH("prologue_bwd_reduce");
  return STY_OK;
}

extern "C" int sty_prologue_bwd_apply(const float* dxp, const float* x, int64_t x_bs, int64_t x_cs,
                                      const float* scale, const float* shift, const float* alpha,
                                      const float* mask, const float* c0, const float* c1, const float* add,
                                      int64_t add_bs, int64_t add_cs, float* dx, int64_t dx_bs, int64_t dx_cs,
                                      int B, int C, int T, int act, sty_stream_t stream) {
  STY_REQUIRE(dxp && x && dx && B > 0 && C > 0 && T > 0, "prologue_bwd_apply: bad argument");
  STY_REQUIRE(act != STY_ACT_SNAKE || alpha, "prologue_bwd_apply: snake needs alpha");
  STY_REQUIRE(cdiv(T, kApplySpan) <= 65535, "prologue_bwd_apply: T too large for the grid");
  dim3 grid((unsigned)((int64_t)B * C), cdiv(T, kApplySpan));
  prologue_bwd_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(dxp, x, x_bs, x_cs, scale, shift, alpha, mask, c0,
                                                                 c1, add, add_bs, add_cs, dx, dx_bs, dx_cs, C, T,
                                                                 act);
  STY_CHECK_LAUNCH("prologue_bwd_apply");
  return STY_OK;
}

extern "C" int sty_grn_snake_bwd(const float* g_u, const float* h, const float* gs, const float* kc,
                                 const float* alpha, float* d_h, float* dalpha, int B, int J, int T,
                                 sty_stream_t stream) {
  STY_REQUIRE(g_u && h && gs && kc && alpha && d_h && dalpha && B > 0 && J > 0 && T > 0, "grn_snake_bwd: bad argument");
  grn_snake_bwd_kernel<<<B * J, 512, 0, as_stream(stream)>>>(g_u, h, gs, kc, alpha, d_h, dalpha, J, T);
  STY_CHECK_LAUNCH("grn_snake_bwd");
  return STY_OK;
}

namespace sty {
namespace {
// generic-activation variant of grn_snake_bwd_kernel: hb = act(h)
__global__ void __launch_bounds__(512)
grn_act_bwd_kernel(const float* g_u, const float* __restrict__ h, const float* __restrict__ gs,
                   const float* __restrict__ kc, float* d_h, int T, int act) {
  const int64_t off = (int64_t)blockIdx.x * T;
  const float s = gs[blockIdx.x], k = kc[blockIdx.x];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float hv = h[off + t];
    const float hb = act_apply(hv, act);
    d_h[off + t] = fmaf(g_u[off + t], s, k * hb) * act_grad(hv, act, 1.f);
  }
}
}  // namespace
}  // namespace sty

extern "C" int sty_grn_act_bwd(const float* g_u, const float* h, const float* gs, const float* kc, const float* alpha,
                               float* d_h, float* dalpha, int B, int J, int T, int act, sty_stream_t stream) {
  if (act == STY_ACT_SNAKE) return sty_grn_snake_bwd(g_u, h, gs, kc, alpha, d_h, dalpha, B, J, T, stream);
  STY_REQUIRE(g_u && h && gs && kc && d_h && B > 0 && J > 0 && T > 0, "grn_act_bwd: bad argument");
  grn_act_bwd_kernel<<<B * J, 512, 0, as_stream(stream)>>>(g_u, h, gs, kc, d_h, T, act);
  STY_CHECK_LAUNCH("grn_act_bwd");
  return STY_OK;
}

extern "C" int sty_chan_layernorm_bwd(const float* x, const float* res, int64_t x_bs, const float* gamma,
                                      const float* beta, int64_t g_bs, int g_plus_one, const float* dy,
                                      const float* mask, float* dv, float* dgb, int64_t dg_bs, int B, int C, int T,
                                      float eps, int act, sty_stream_t stream) {
  STY_REQUIRE(x && gamma && beta && dy && dv && dgb && B > 0 && C > 0 && T > 0, "chan_layernorm_bwd: bad argument");
  STY_REQUIRE(act == STY_ACT_NONE || act == STY_ACT_RELU, "chan_layernorm_bwd: only none/relu epilogues");
  STY_REQUIRE((2 * (size_t)C + 4 * 2 * 32 * 33) * sizeof(float) <= 48 * 1024,
              "chan_layernorm_bwd: C=%d needs more than 48 KB of shared memory", C);
  dim3 grid(cdiv(T, 128), B);
  chan_layernorm_bwd_kernel<<<grid, 128, (2 * C + 4 * 2 * 32 * 33) * sizeof(float), as_stream(stream)>>>(
      x, res, x_bs, gamma, beta, g_bs, g_plus_one, dy, mask, dv, dgb, dg_bs, C, T, eps, act);
  STY_CHECK_LAUNCH("chan_layernorm_bwd");
  return STY_OK;
}

extern "C" int sty_dwconv1d_bwd(const float* dy, const float* x, int64_t x_bs, int64_t x_cs, const float* w,
                                float* dx, int64_t dx_bs, int64_t dx_cs, float* dw, float* db, int B, int C, int T,
                                int K, int pad_left, sty_stream_t stream) {
  STY_REQUIRE(dy && x && w && dw && B > 0 && C > 0 && T > 0 && K >= 1 && K <= 31, "dwconv1d_bwd: bad argument (K<=31)");
  STY_REQUIRE(B <= 65535, "dwconv1d_bwd: batch too large");
  dim3 grid(C, B);
  cudaStream_t st = as_stream(stream);
  if (K == 7 && pad_left == 3)
    dwconv1d_bwd_win_kernel<7, 3, 4><<<grid, 256, 0, st>>>(dy, x, x_bs, x_cs, w, dx, dx_bs, dx_cs, dw, db, C, T);
  else if (K <= 7)
    dwconv1d_bwd_kernel<7><<<grid, 256, 0, st>>>(dy, x, x_bs, x_cs, w, dx, dx_bs, dx_cs, dw, db, C, T, K, pad_left);
  else
    dwconv1d_bwd_kernel<31><<<grid, 256, 0, st>>>(dy, x, x_bs, x_cs, w, dx, dx_bs, dx_cs, dw, db, C, T, K, pad_left);
  STY_CHECK_LAUNCH("dwconv1d_bwd");
  return STY_OK;
}

extern "C" int sty_glu_bwd(const float* x, const float* dy, float* dx, int B, int C, int T, sty_stream_t stream) {
  STY_REQUIRE(x && dy && dx && B > 0 && C > 0 && T > 0, "glu_bwd: bad argument");
  dim3 grid(cdiv((int64_t)C * T, 256), B);
  glu_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dy, dx, C, T);
  STY_CHECK_LAUNCH("glu_bwd");
  return STY_OK;
}

extern "C" int sty_unshuffle(const float* y, float* x, int B, int CO, int T, int s, sty_stream_t stream) {
  STY_REQUIRE(y && x && B > 0 && CO > 0 && T > 0 && s > 1 && CO % s == 0 && CO <= 65535, "unshuffle: bad argument");
  dim3 grid(cdiv(T, 256), CO, B);
  unshuffle_kernel<<<grid, 256, 0, as_stream(stream)>>>(y, x, CO, T, s);
  STY_CHECK_LAUNCH("unshuffle");
  return STY_OK;
}

extern "C" int sty_embed_bwd(const int64_t* tokens, const int64_t* lengths, const float* dx, float* d_emb, int B,
                             int T, int C, int n_tokens, float scale, sty_stream_t stream) {
  STY_REQUIRE(tokens && dx && d_emb && B > 0 && T > 0 && C > 0 && n_tokens > 0, "embed_bwd: bad argument");
  dim3 grid(cdiv(T, 256), C, B);
  embed_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(tokens, lengths, dx, d_emb, T, C, n_tokens, scale);
  STY_CHECK_LAUNCH("embed_bwd");
  return STY_OK;
}

extern "C" int sty_bmm_nt_fwd(const float* A, int64_t a_bs, const float* Bt, int64_t b_bs, float* C, int64_t c_bs,
                              int B, int M, int N, int K, sty_stream_t stream) {
  STY_REQUIRE(A && Bt && C && B > 0 && M > 0 && N > 0 && K > 0, "bmm_nt: bad argument");
  dim3 grid(cdiv(N, 64), cdiv(M, 64), B);
  bmm_nt_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, a_bs, Bt, b_bs, C, c_bs, M, N, K);
  STY_CHECK_LAUNCH("bmm_nt");
  return STY_OK;
}

extern "C" int sty_linear_rows_bwd(const float* dh, const float* s, const float* W, float* dW, float* dbias,
                                   float* ds, int B, int I, int J, sty_stream_t stream) {
  STY_REQUIRE(dh && s && W && dW && dbias && B > 0 && I > 0 && I <= 256 && J > 0, "linear_rows_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  linear_rows_bwd_w_kernel<<<cdiv((int64_t)J * I, 256), 256, 0, st>>>(dh, s, dW, dbias, B, I, J);
  STY_CHECK_LAUNCH("linear_rows_bwd_w");
  if (ds) {
    linear_rows_bwd_s_kernel<<<B, 256, 8 * I * sizeof(float), st>>>(dh, W, ds, I, J);
    STY_CHECK_LAUNCH("linear_rows_bwd_s");
  }
  return STY_OK;
}

extern "C" int sty_istft_head_bwd(const float* dout, const float* out, const float* logamp, int64_t logamp_bs,
                                  const float* real, const float* imag, int64_t ri_bs, const float* basis_re,
                                  const float* basis_im, float* d_logamp, float* d_real, float* d_imag,
                                  int64_t dri_bs, int B, int S, int bins, int n_fft, int hop,
                                  sty_stream_t stream) {
  STY_REQUIRE(dout && out && logamp && real && imag && basis_re && basis_im && d_logamp && d_real && d_imag,
              "istft_head_bwd: null pointer");
  STY_REQUIRE(n_fft == 64 && hop == 4 && bins == 32, "istft_head_bwd: built for n_fft=64 hop=4 bins=32");
  STY_REQUIRE(B > 0 && S > 1, "istft_head_bwd: bad shape");
  constexpr int FT = 128;
  const size_t smem = ((size_t)2 * 32 * 64 + FT * 4 + 64) * sizeof(float);
  auto kern = istft_head_bwd_kernel<64, 4, 32, FT>;
  cudaStream_t st = as_stream(stream);
  // column S-1 receives two atomic contributions (frames S-1 and S): zero it first
  cudaMemset2DAsync(d_logamp + (S - 1), (size_t)S * sizeof(float), 0, sizeof(float), (size_t)B * bins, st);
  for (int b = 0; b < B; ++b)
    for (float* d : {d_real, d_imag})
      cudaMemset2DAsync(d + (int64_t)b * dri_bs + (S - 1), (size_t)S * sizeof(float), 0, sizeof(float), (size_t)bins,
                        st);
  dim3 grid(cdiv(S + 1, FT), B);
  kern<<<grid, FT, smem, st>>>(dout, out, logamp, logamp_bs, real, imag, ri_bs, basis_re, basis_im, d_logamp,
                               d_real, d_imag, dri_bs, S);
  STY_CHECK_LAUNCH("istft_head_bwd");
  return STY_OK;
}

extern "C" int sty_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float grad_scale,
                              sty_stream_t stream) {
  STY_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw_step: bad argument");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 1184) blocks = 1184;
  adamw_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                               bc1, bc2, grad_scale);
  STY_CHECK_LAUNCH("adamw_step");
  return STY_OK;
}

extern "C" int sty_adamw_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                                  float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                  sty_stream_t stream) {
  STY_REQUIRE(p && g && m && v && hyper && n > 0, "adamw_step_dev: bad argument");
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 1184) blocks = 1184;
  adamw_dev_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, hyper, beta1, beta2, eps,
                                                                   weight_decay, grad_scale);
  STY_CHECK_LAUNCH("adamw_step_dev");
  return STY_OK;
}
