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
                   float* __restrict__ dq, int64_t dq_bs, float* __restrict__ delta) {
  constexpr int DP = D + 4;
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
      const float ds = expf(s - ls) * (dp - di);
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[u * DP + j]);
        acc[j] = fmaf(ds, kk.x, acc[j]); acc[j + 1] = fmaf(ds, kk.y, acc[j + 1]);
        acc[j + 2] = fmaf(ds, kk.z, acc[j + 2]); acc[j + 3] = fmaf(ds, kk.w, acc[j + 3]);
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
                    float* __restrict__ dk, float* __restrict__ dv, int64_t dq_bs) {
  constexpr int DP = D + 4;
  constexpr bool DK = MODE != 1, DV = MODE != 2;
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
      if constexpr (DV) {
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 gg = *reinterpret_cast<const float4*>(&Gs[u * DP + j]);
          av[j] = fmaf(p, gg.x, av[j]); av[j + 1] = fmaf(p, gg.y, av[j + 1]);
          av[j + 2] = fmaf(p, gg.z, av[j + 2]); av[j + 3] = fmaf(p, gg.w, av[j + 3]);
        }
      }
      if constexpr (DK) {
        const float ds = p * (dp - Ds[u]);
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 qq = *reinterpret_cast<const float4*>(&Qs[u * DP + j]);
          ak[j] = fmaf(ds, qq.x, ak[j]); ak[j + 1] = fmaf(ds, qq.y, ak[j + 1]);
          ak[j + 2] = fmaf(ds, qq.z, ak[j + 2]); ak[j + 3] = fmaf(ds, qq.w, ak[j + 3]);
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

extern "C" int sty_attention_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                                 const float* d_o, int64_t o_bs, const float* lse, const int64_t* lengths,
                                 const float* rope_cos, const float* rope_sin, int d_rot, float* dq, float* dk,
                                 float* dv, int64_t dqkv_bs, float* delta, int B, int H, int D, int T, float scale,
                                 sty_stream_t stream) {
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
                                                             rope_sin, T, scale, dq, dqkv_bs, delta);
      attn_bwd_dkv_kernel<16, 32, QB, 4, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                                 rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs);
    } else {
      attn_bwd_dq_kernel<16, 32, QB, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos,
                                                             rope_sin, T, scale, dq, dqkv_bs, delta);
      attn_bwd_dkv_kernel<16, 32, QB, 0, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                                 rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs);
    }
  } else if (D == 64) {
    STY_REQUIRE(!rope_cos, "attention_bwd: D=64 is built without RoPE");
    constexpr int QB = 128;
    dim3 grid(cdiv(T, QB), H, B);
    attn_bwd_dq_kernel<64, 16, QB, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, d_o, o_bs, lse, lengths, rope_cos,
                                                           rope_sin, T, scale, dq, dqkv_bs, delta);
    attn_bwd_dkv_kernel<64, 16, QB, 0, 1><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                               rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs);
    attn_bwd_dkv_kernel<64, 16, QB, 0, 2><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, d_o, o_bs, lse, delta, lengths,
                                                               rope_cos, rope_sin, T, scale, dk, dv, dqkv_bs);
  } else {
    set_error("attention_bwd: unsupported head dim %d (built: 16, 64)", D);
    return STY_ERR_BAD_ARG;
  }
  STY_CHECK_LAUNCH("attention_bwd");
  return STY_OK;
}
