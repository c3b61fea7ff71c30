// Backward of the multi-head attention core (fp32): recomputes the probabilities
// from the saved log-sum-exp instead of storing the (T x T) matrices.
//   dq kernel : one thread per query row, K/V tiles streamed through shared memory;
//               also writes delta[b,h,t] = <dO_t, O_t>.
//   dkv kernel: one thread per key row, Q/dO tiles streamed through shared memory.
// RoPE is re-applied to q/k on load and its transpose to dq/dk on store; the additive
// -1e4 padding mask is the forward's.  Layout as in attention.cu: (B, H*D, T).
#include <math.h>

#include "common.cuh"

namespace sty {
namespace {

template <int D, int HALF>
__device__ __forceinline__ void rope_fwd(float* r, const float* __restrict__ c, const float* __restrict__ s, int t) {
  if constexpr (HALF > 0) {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      const float cc = c[(int64_t)t * HALF + i], ss = s[(int64_t)t * HALF + i];
      const float a = r[i], b = r[i + HALF];
      r[i] = a * cc - b * ss;
      r[i + HALF] = b * cc + a * ss;
    }
  }
}

template <int D, int HALF>
__device__ __forceinline__ void rope_bwd(float* r, const float* __restrict__ c, const float* __restrict__ s, int t) {
  if constexpr (HALF > 0) {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
      const float cc = c[(int64_t)t * HALF + i], ss = s[(int64_t)t * HALF + i];
      const float a = r[i], b = r[i + HALF];
      r[i] = a * cc + b * ss;
      r[i + HALF] = b * cc - a * ss;
    }
  }
}

template <int D, int KT, int QB, int HALF>
__global__ void __launch_bounds__(QB)
attn_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                   int64_t qkv_bs, const float* __restrict__ o, const float* __restrict__ dO, int64_t o_bs,
                   const float* __restrict__ lse, const int64_t* __restrict__ lengths,
                   const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int T, float scale,
                   float* __restrict__ dq, int64_t dq_bs, float* __restrict__ delta, DropSpec ds) {
  constexpr int DP = D + 4;
  const unsigned long long seed = ds.seed ? *ds.seed : 0ull;
  const unsigned long long drow = (((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * T +
                                   (blockIdx.x * QB + threadIdx.x)) * (unsigned long long)T;
  __shared__ __align__(16) float Ks[KT * DP];
  __shared__ __align__(16) float Vs[KT * DP];
  const int b = blockIdx.z, h = blockIdx.y, tid = threadIdx.x;
  const int tq = blockIdx.x * QB + tid;
  const bool q_ok = tq < T;
  const int len = lengths ? (int)lengths[b] : T;
  const bool q_valid = tq < len;
  const int64_t hoff = (int64_t)h * D * T;
  const float* __restrict__ qb = q + (int64_t)b * qkv_bs + hoff;
  const float* __restrict__ kb = k + (int64_t)b * qkv_bs + hoff;
  const float* __restrict__ vb = v + (int64_t)b * qkv_bs + hoff;
  const float* __restrict__ ob = o + (int64_t)b * o_bs + hoff;
  const float* __restrict__ dob = dO + (int64_t)b * o_bs + hoff;
  float qr[D], dor[D], acc[D];
  float di = 0.f;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    qr[j] = q_ok ? qb[(int64_t)j * T + tq] : 0.f;
    dor[j] = q_ok ? dob[(int64_t)j * T + tq] : 0.f;
    const float ov = q_ok ? ob[(int64_t)j * T + tq] : 0.f;
    di = fmaf(dor[j], ov, di);
    acc[j] = 0.f;
  }
  if (q_ok) rope_fwd<D, HALF>(qr, rope_cos, rope_sin, tq);
#pragma unroll
  for (int j = 0; j < D; ++j) qr[j] *= scale;
  const int64_t row = ((int64_t)b * gridDim.y + h) * T + tq;
  const float ls = q_ok ? lse[row] : 0.f;
  if (q_ok) delta[row] = di;

  for (int k0 = 0; k0 < T; k0 += KT) {
    for (int idx = tid; idx < KT * D; idx += QB) {
      const int j = idx / KT, u = idx - j * KT;
      const int t = k0 + u;
      Ks[u * DP + j] = t < T ? kb[(int64_t)j * T + t] : 0.f;
      Vs[u * DP + j] = t < T ? vb[(int64_t)j * T + t] : 0.f;
    }
    __syncthreads();
    if constexpr (HALF > 0) {
      for (int u = tid; u < KT; u += QB)
        if (k0 + u < T) rope_fwd<D, HALF>(&Ks[u * DP], rope_cos, rope_sin, k0 + u);
      __syncthreads();
    }
    const int nk = min(KT, T - k0);
    for (int u = 0; u < nk; ++u) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[u * DP + j]);
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[u * DP + j]);
        s = fmaf(qr[j], kk.x, s); s = fmaf(qr[j + 1], kk.y, s);
        s = fmaf(qr[j + 2], kk.z, s); s = fmaf(qr[j + 3], kk.w, s);
        dp = fmaf(dor[j], vv.x, dp); dp = fmaf(dor[j + 1], vv.y, dp);
        dp = fmaf(dor[j + 2], vv.z, dp); dp = fmaf(dor[j + 3], vv.w, dp);
      }
      if (lengths && !(q_valid && (k0 + u) < len)) s += -1e4f;
      if (ds.seed) dp = drop_keep(seed, ds.site, drow + (unsigned long long)(k0 + u), ds.thresh) ? dp * ds.inv_keep : 0.f;
      const float dsc = expf(s - ls) * (dp - di);
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[u * DP + j]);
        acc[j] = fmaf(dsc, kk.x, acc[j]); acc[j + 1] = fmaf(dsc, kk.y, acc[j + 1]);
        acc[j + 2] = fmaf(dsc, kk.z, acc[j + 2]); acc[j + 3] = fmaf(dsc, kk.w, acc[j + 3]);
      }
    }
    __syncthreads();
  }
  if (q_ok) {
#pragma unroll
    for (int j = 0; j < D; ++j) acc[j] *= scale;
    rope_bwd<D, HALF>(acc, rope_cos, rope_sin, tq);
    float* __restrict__ dqb = dq + (int64_t)b * dq_bs + hoff;
#pragma unroll
    for (int j = 0; j < D; ++j) dqb[(int64_t)j * T + tq] = acc[j];
  }
}

// MODE 0: dk and dv; 1: dv only; 2: dk only (register budget at D = 64)
template <int D, int QT, int KB, int HALF, int MODE>
__global__ void __launch_bounds__(KB)
attn_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                    int64_t qkv_bs, const float* __restrict__ dO, int64_t o_bs, const float* __restrict__ lse,
                    const float* __restrict__ delta, const int64_t* __restrict__ lengths,
                    const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int T, float scale,
                    float* __restrict__ dk, float* __restrict__ dv, int64_t dq_bs, DropSpec ds) {
  constexpr int DP = D + 4;
  constexpr bool DK = MODE != 1, DV = MODE != 2;
  const unsigned long long seed = ds.seed ? *ds.seed : 0ull;
  const unsigned long long dbase = ((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * T;
  __shared__ __align__(16) float Qs[QT * DP];
  __shared__ __align__(16) float Gs[QT * DP];
  __shared__ float Ls[QT], Ds[QT];
  const int b = blockIdx.z, h = blockIdx.y, tid = threadIdx.x;
  const int tk = blockIdx.x * KB + tid;
  const bool k_ok = tk < T;
  const int len = lengths ? (int)lengths[b] : T;
  const bool k_valid = tk < len;
  const int64_t hoff = (int64_t)h * D * T;
  const float* __restrict__ qb = q + (int64_t)b * qkv_bs + hoff;
  const float* __restrict__ kb = k + (int64_t)b * qkv_bs + hoff;
  const float* __restrict__ vb = v + (int64_t)b * qkv_bs + hoff;
  const float* __restrict__ dob = dO + (int64_t)b * o_bs + hoff;
  const int64_t rbase = ((int64_t)b * gridDim.y + h) * T;
  float kr[D], vr[DK ? D : 1], ak[DK ? D : 1], av[DV ? D : 1];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    kr[j] = k_ok ? kb[(int64_t)j * T + tk] : 0.f;
    if constexpr (DK) {
      vr[j] = k_ok ? vb[(int64_t)j * T + tk] : 0.f;
      ak[j] = 0.f;
    }
    if constexpr (DV) av[j] = 0.f;
  }
  if (k_ok) rope_fwd<D, HALF>(kr, rope_cos, rope_sin, tk);

  for (int q0 = 0; q0 < T; q0 += QT) {
    for (int idx = tid; idx < QT * D; idx += KB) {
      const int j = idx / QT, u = idx - j * QT;
      const int t = q0 + u;
      Qs[u * DP + j] = t < T ? qb[(int64_t)j * T + t] : 0.f;
      Gs[u * DP + j] = t < T ? dob[(int64_t)j * T + t] : 0.f;
    }
    for (int u = tid; u < QT; u += KB) {
      Ls[u] = q0 + u < T ? lse[rbase + q0 + u] : 0.f;
      Ds[u] = q0 + u < T ? delta[rbase + q0 + u] : 0.f;
    }
    __syncthreads();
    for (int u = tid; u < QT; u += KB) {
      if (q0 + u < T) rope_fwd<D, HALF>(&Qs[u * DP], rope_cos, rope_sin, q0 + u);
#pragma unroll
      for (int j = 0; j < D; ++j) Qs[u * DP + j] *= scale;
    }
    __syncthreads();
    const int nq = min(QT, T - q0);
    for (int u = 0; u < nq; ++u) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 qq = *reinterpret_cast<const float4*>(&Qs[u * DP + j]);
        s = fmaf(qq.x, kr[j], s); s = fmaf(qq.y, kr[j + 1], s);
        s = fmaf(qq.z, kr[j + 2], s); s = fmaf(qq.w, kr[j + 3], s);
        if constexpr (DK) {
          const float4 gg = *reinterpret_cast<const float4*>(&Gs[u * DP + j]);
          dp = fmaf(gg.x, vr[j], dp); dp = fmaf(gg.y, vr[j + 1], dp);
          dp = fmaf(gg.z, vr[j + 2], dp); dp = fmaf(gg.w, vr[j + 3], dp);
        }
      }
      if (lengths && !(k_valid && (q0 + u) < len)) s += -1e4f;
      const float p = expf(s - Ls[u]);
      float km = 1.f;  // keep mask * 1/(1-p) of probability (query q0+u, key tk)
      if (ds.seed)
        km = drop_keep(seed, ds.site, (dbase + (unsigned long long)(q0 + u)) * (unsigned long long)T + tk, ds.thresh)
                 ? ds.inv_keep : 0.f;
      if constexpr (DV) {
        const float pd = p * km;
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 gg = *reinterpret_cast<const float4*>(&Gs[u * DP + j]);
          av[j] = fmaf(pd, gg.x, av[j]); av[j + 1] = fmaf(pd, gg.y, av[j + 1]);
          av[j + 2] = fmaf(pd, gg.z, av[j + 2]); av[j + 3] = fmaf(pd, gg.w, av[j + 3]);
        }
      }
      if constexpr (DK) {
        const float dsc = p * (dp * km - Ds[u]);
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 qq = *reinterpret_cast<const float4*>(&Qs[u * DP + j]);
          ak[j] = fmaf(dsc, qq.x, ak[j]); ak[j + 1] = fmaf(dsc, qq.y, ak[j + 1]);
          ak[j + 2] = fmaf(dsc, qq.z, ak[j + 2]); ak[j + 3] = fmaf(dsc, qq.w, ak[j + 3]);
        }
      }
    }
    __syncthreads();
  }
  if (k_ok) {
    if constexpr (DK) {
      rope_bwd<D, HALF>(ak, rope_cos, rope_sin, tk);
      float* __restrict__ d = dk + (int64_t)b * dq_bs + hoff;
#pragma unroll
      for (int j = 0; j < D; ++j) d[(int64_t)j * T + tk] = ak[j];
    }
    if constexpr (DV) {
      float* __restrict__ d = dv + (int64_t)b * dq_bs + hoff;
#pragma unroll
      for (int j = 0; j < D; ++j) d[(int64_t)j * T + tk] = av[j];
    }
  }
}

}  // namespace
}  // namespace sty

using namespace sty;

static int attention_bwd_launch(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                                const float* d_o, int64_t o_bs, const float* lse, const int64_t* lengths,
                                const float* rope_cos, const float* rope_sin, int d_rot, float* dq, float* dk,
                                float* dv, int64_t dqkv_bs, float* delta, int B, int H, int D, int T, float scale,
                                const sty_dropout* drop, sty_stream_t stream) {
  STY_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "attention_bwd: dropout p must be in [0,1)");
  const DropSpec ds = make_drop(drop);
  STY_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv && delta, "attention_bwd: null pointer");
  STY_REQUIRE(B > 0 && H > 0 && T > 0 && H <= 65535 && B <= 65535, "attention_bwd: bad shape");
  STY_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "attention_bwd: need both rope tables");
  cudaStream_t st = as_stream(stream);
  if (D == 16) {
    constexpr int QB = 64;
    dim3 grid(cdiv(T, QB), H, B);
    if (rope_cos) {
      STY_REQUIRE(d_rot == 8, "attention_bwd: D=16 is built with d_rot=8 (got %d)", d_rot);
      attn_bwd_dq_kernel<16, 32, QB, 4><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos,
                                                             rope_sin, T, scale, dq, dqkv_bs, delta, ds);
      attn_bwd_dkv_kernel<16, 32, QB, 4, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                                 rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs, ds);
    } else {
      attn_bwd_dq_kernel<16, 32, QB, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos,
                                                             rope_sin, T, scale, dq, dqkv_bs, delta, ds);
      attn_bwd_dkv_kernel<16, 32, QB, 0, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                                 rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs, ds);
    }
  } else if (D == 64) {
    STY_REQUIRE(!rope_cos, "attention_bwd: D=64 is built without RoPE");
    constexpr int QB = 128;
    dim3 grid(cdiv(T, QB), H, B);
    attn_bwd_dq_kernel<64, 16, QB, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos,
                                                           rope_sin, T, scale, dq, dqkv_bs, delta, ds);
    attn_bwd_dkv_kernel<64, 16, QB, 0, 1><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                               rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs, ds);
    attn_bwd_dkv_kernel<64, 16, QB, 0, 2><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                               rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs, ds);
  } else {
    set_error("attention_bwd: unsupported head dim %d (built: 16, 64)", D);
    return STY_ERR_BAD_ARG;
  }
  STY_CHECK_LAUNCH("attention_bwd");
  return STY_OK;
}

extern "C" int sty_attention_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                                 const float* d_o, int64_t o_bs, const float* lse, const int64_t* lengths,
                                 const float* rope_cos, const float* rope_sin, int d_rot, float* dq, float* dk,
                                 float* dv, int64_t dqkv_bs, float* delta, int B, int H, int D, int T, float scale,
                                 sty_stream_t stream) {
  return attention_bwd_launch(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos, rope_sin, d_rot, dq, dk, dv,
                              dqkv_bs, delta, B, H, D, T, scale, nullptr, stream);
}

extern "C" int sty_attention_drop_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs,
                                      const float* o, const float* d_o, int64_t o_bs, const float* lse,
                                      const int64_t* lengths, const float* rope_cos, const float* rope_sin,
                                      int d_rot, float* dq, float* dk, float* dv, int64_t dqkv_bs, float* delta,
                                      int B, int H, int D, int T, float scale, const sty_dropout* drop,
                                      sty_stream_t stream) {
  return attention_bwd_launch(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos, rope_sin, d_rot, dq, dk, dv,
                              dqkv_bs, delta, B, H, D, T, scale, drop, stream);
}

// =====================================================================================================
// Generic head size (prosody encoder: 2 heads x 160): backward through MATERIALISED probabilities —
// T is a few hundred tokens, so P (B,H,T,T) is a few MB.  Pieces: heads <-> token-major rows (with RoPE
// and its transpose), row softmax of the scores, softmax backward, and the batched products
// (sty_bmm_fwd / sty_bmm_nt_fwd / sty_bmm_tn_fwd).
// =====================================================================================================
namespace sty {
namespace {

// x (B, H*D, T) [batch stride x_bs] -> y (B, H, T, D) * mul, RoPE on the first d_rot features (rope != null)
__global__ void __launch_bounds__(256)
heads_to_rows_kernel(const float* __restrict__ x, int64_t x_bs, float* __restrict__ y, const float* __restrict__ rc,
                     const float* __restrict__ rs, int half, int H, int D, int T, float mul) {
  __shared__ float tile[32][33];
  const int bh = blockIdx.z, b = bh / H, h = bh - b * H;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* __restrict__ xb = x + (int64_t)b * x_bs + (int64_t)h * D * T;
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, t = t0 + tx;
    float v = 0.f;
    if (d < D && t < T) {
      v = xb[(int64_t)d * T + t];
      if (rc && d < 2 * half) {  // rotate-half convention on the first d_rot = 2*half features
        const int i = d < half ? d : d - half;
        const float c = rc[(int64_t)t * half + i], s = rs[(int64_t)t * half + i];
        const float other = xb[(int64_t)(d < half ? d + half : d - half) * T + t];
        v = d < half ? v * c - other * s : v * c + other * s;
      }
    }
    tile[r][tx] = v * mul;
  }
  __syncthreads();
  float* __restrict__ yb = y + (int64_t)bh * T * D;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, d = d0 + tx;
    if (t < T && d < D) yb[(int64_t)t * D + d] = tile[tx][r];
  }
}

// y (B,H,T,D) -> x (B, H*D, T) [batch stride x_bs] * mul, with the TRANSPOSE of the RoPE rotation
__global__ void __launch_bounds__(256)
rows_to_heads_kernel(const float* __restrict__ y, float* __restrict__ x, int64_t x_bs, const float* __restrict__ rc,
                     const float* __restrict__ rs, int half, int H, int D, int T, float mul) {
  __shared__ float tile[32][33];
  const int bh = blockIdx.z, b = bh / H, h = bh - b * H;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* __restrict__ yb = y + (int64_t)bh * T * D;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, d = d0 + tx;
    float v = 0.f;
    if (t < T && d < D) {
      v = yb[(int64_t)t * D + d];
      if (rc && d < 2 * half) {
        const int i = d < half ? d : d - half;
        const float c = rc[(int64_t)t * half + i], s = rs[(int64_t)t * half + i];
        const float other = yb[(int64_t)t * D + (d < half ? d + half : d - half)];
        v = d < half ? v * c + other * s : v * c - other * s;
      }
    }
    tile[r][tx] = v * mul;
  }
  __syncthreads();
  float* __restrict__ xb = x + (int64_t)b * x_bs + (int64_t)h * D * T;
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, t = t0 + tx;
    if (d < D && t < T) xb[(int64_t)d * T + t] = tile[tx][r];
  }
}

// P[bh,i,:] = softmax_j( <q_r[i], k_r[j]> + mask(i,j) ), q_r already scaled; one CTA per (bh, i)
__global__ void __launch_bounds__(128)
attn_probs_kernel(const float* __restrict__ qr, const float* __restrict__ kr, const int64_t* __restrict__ lengths,
                  float* __restrict__ P, int H, int D, int T) {
  extern __shared__ float sm[];  // q row [D] + scores [T]
  __shared__ float red[32];
  float* qs = sm;
  float* sc = sm + D;
  const int bh = blockIdx.y, i = blockIdx.x, b = bh / H;
  const int len = lengths ? (int)lengths[b] : T;
  const float* __restrict__ qrow = qr + ((int64_t)bh * T + i) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) qs[d] = qrow[d];
  __syncthreads();
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    const float* __restrict__ krow = kr + ((int64_t)bh * T + j) * D;
    float a = 0.f;
    for (int d = 0; d < D; ++d) a = fmaf(qs[d], krow[d], a);
    if (lengths && !(i < len && j < len)) a += -1e4f;
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  // block max / sum
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  float s = 0.f;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    s += e;
  }
  s = block_sum(s, red);
  const float inv = 1.f / s;
  float* __restrict__ prow = P + ((int64_t)bh * T + i) * T;
  for (int j = threadIdx.x; j < T; j += blockDim.x) prow[j] = sc[j] * inv;
}

// dS = P * (dP - sum_j P dP), in place on dP; one CTA per row
__global__ void __launch_bounds__(128)
softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ dP, int T) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* __restrict__ p = P + row * T;
  float* __restrict__ g = dP + row * T;
  float s = 0.f;
  for (int j = threadIdx.x; j < T; j += blockDim.x) s = fmaf(p[j], g[j], s);
  s = block_sum(s, red);
  for (int j = threadIdx.x; j < T; j += blockDim.x) g[j] = p[j] * (g[j] - s);
}

// C[b] (M,N) = A[b] (K,M)^T @ Bm[b] (K,N)
__global__ void __launch_bounds__(256)
bmm_tn_kernel(const float* __restrict__ A, int64_t a_bs, const float* __restrict__ Bm, int64_t b_bs,
              float* __restrict__ Cm, int64_t c_bs, int M, int N, int K) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const float* __restrict__ Ab = A + (int64_t)b * a_bs;
  const float* __restrict__ Bb = Bm + (int64_t)b * b_bs;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int idx = tid; idx < BK * BM; idx += 256) {
      const int mm = idx % BM, kk = idx / BM;
      const int k = k0 + kk;
      As[kk][mm] = (m0 + mm < M && k < K) ? Ab[(int64_t)k * M + m0 + mm] : 0.f;
      Bs[kk][mm] = (n0 + mm < N && k < K) ? Bb[(int64_t)k * N + n0 + mm] : 0.f;
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

}  // namespace
}  // namespace sty

extern "C" int sty_heads_to_rows(const float* x, int64_t x_bs, float* y, const float* rope_cos, const float* rope_sin,
                                 int d_rot, int B, int H, int D, int T, float mul, int inverse, sty_stream_t stream) {
  STY_REQUIRE(x && y && B > 0 && H > 0 && D > 0 && T > 0, "heads_to_rows: bad argument");
  STY_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr) && d_rot % 2 == 0 && d_rot <= D, "heads_to_rows: bad rope");
  dim3 grid(cdiv(T, 32), cdiv(D, 32), B * H);
  if (!inverse)
    heads_to_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, x_bs, y, rope_cos, rope_sin, d_rot / 2, H, D, T, mul);
  else  // x is the (B,H,T,D) input, y the (B,H*D,T) output with batch stride x_bs
    rows_to_heads_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, x_bs, rope_cos, rope_sin, d_rot / 2, H, D, T, mul);
  STY_CHECK_LAUNCH("heads_to_rows");
  return STY_OK;
}

extern "C" int sty_attn_probs(const float* q_rows, const float* k_rows, const int64_t* lengths, float* P, int B, int H,
                              int D, int T, sty_stream_t stream) {
  STY_REQUIRE(q_rows && k_rows && P && B > 0 && H > 0 && D > 0 && T > 0 && B * H <= 65535, "attn_probs: bad argument");
  dim3 grid(T, B * H);
  attn_probs_kernel<<<grid, 128, (size_t)(D + T) * sizeof(float), as_stream(stream)>>>(q_rows, k_rows, lengths, P, H, D, T);
  STY_CHECK_LAUNCH("attn_probs");
  return STY_OK;
}

extern "C" int sty_softmax_bwd(const float* P, float* dP, int64_t rows, int T, sty_stream_t stream) {
  STY_REQUIRE(P && dP && rows > 0 && T > 0, "softmax_bwd: bad argument");
  softmax_bwd_kernel<<<(unsigned)rows, 128, 0, as_stream(stream)>>>(P, dP, T);
  STY_CHECK_LAUNCH("softmax_bwd");
  return STY_OK;
}

extern "C" int sty_bmm_tn_fwd(const float* A, int64_t a_bs, const float* Bm, int64_t b_bs, float* C, int64_t c_bs,
                              int B, int M, int N, int K, sty_stream_t stream) {
  STY_REQUIRE(A && Bm && C && B > 0 && M > 0 && N > 0 && K > 0, "bmm_tn: bad argument");
  dim3 grid(cdiv(N, 64), cdiv(M, 64), B);
  bmm_tn_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, a_bs, Bm, b_bs, C, c_bs, M, N, K);
  STY_CHECK_LAUNCH("bmm_tn");
  return STY_OK;
}
