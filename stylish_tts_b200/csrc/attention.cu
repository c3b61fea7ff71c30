// Multi-head attention core (fp32, flash-style streaming softmax) with RoPE and
// the reference's additive -1e4 padding mask fused in.
//
// Layout: q,k,v,o are channel-major (B, H*D, T); head h owns channels
// [h*D,(h+1)*D).  One thread owns one query row (q, running max/sum and the
// D-wide accumulator live in registers); K/V tiles of KT keys are staged in
// shared memory transposed to [key][d] so every thread reads them as broadcast
// 128-bit loads.
//
// This is the fp32 SIMT path (parity-first).  See DESIGN.md for the tcgen05 plan.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace sty {

template <int D, int KT, int QB, int HALF>
__global__ void __launch_bounds__(QB)
attention_kernel(const float* __restrict__ q, const float* __restrict__ k,
                 const float* __restrict__ v, int64_t qkv_bs, float* __restrict__ o, int64_t o_bs,
                 const int64_t* __restrict__ lengths, const float* __restrict__ rope_cos,
                 const float* __restrict__ rope_sin, int T, float scale, float* __restrict__ lse, DropSpec ds) {
  constexpr int DP = D + 4;  // padded row (keeps 16 B alignment, spreads banks)
  const unsigned long long seed = ds.seed ? *ds.seed : 0ull;
  const unsigned long long drow = (((unsigned long long)blockIdx.z * gridDim.y + blockIdx.y) * T +
                                   (blockIdx.x * QB + threadIdx.x)) * (unsigned long long)T;
  __shared__ __align__(16) float Ks[KT * DP];
  __shared__ __align__(16) float Vs[KT * DP];
  const int b = blockIdx.z, h = blockIdx.y;
  const int tid = threadIdx.x;
  const int tq = blockIdx.x * QB + tid;
  const bool q_ok = tq < T;
  constexpr int half = HALF;
  const int len = lengths ? (int)lengths[b] : T;
  const bool q_valid = tq < len;
  const float* __restrict__ qb = q + (int64_t)b * qkv_bs + (int64_t)h * D * T;
  const float* __restrict__ kb = k + (int64_t)b * qkv_bs + (int64_t)h * D * T;
  const float* __restrict__ vb = v + (int64_t)b * qkv_bs + (int64_t)h * D * T;

  float qr[D];
#pragma unroll
  for (int j = 0; j < D; ++j) qr[j] = q_ok ? qb[(int64_t)j * T + tq] : 0.f;
  if constexpr (HALF > 0) {
    if (q_ok) {
#pragma unroll
      for (int i = 0; i < HALF; ++i) {
        const float c = rope_cos[(int64_t)tq * HALF + i], s = rope_sin[(int64_t)tq * HALF + i];
        const float a = qr[i], bb = qr[i + HALF];
        qr[i] = a * c - bb * s;
        qr[i + HALF] = bb * c + a * s;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < D; ++j) qr[j] *= scale;

  float m = -INFINITY, l = 0.f;
  float acc[D];
#pragma unroll
  for (int j = 0; j < D; ++j) acc[j] = 0.f;

  for (int k0 = 0; k0 < T; k0 += KT) {
    // stage K/V tile transposed: global reads coalesced along keys
    for (int idx = tid; idx < KT * D; idx += QB) {
      const int j = idx / KT, u = idx - j * KT;
      const int t = k0 + u;
      float kv = 0.f, vv = 0.f;
      if (t < T) {
        kv = kb[(int64_t)j * T + t];
        vv = vb[(int64_t)j * T + t];
      }
      Ks[u * DP + j] = kv;
      Vs[u * DP + j] = vv;
    }
    __syncthreads();
    if constexpr (HALF > 0) {
      for (int idx = tid; idx < KT * half; idx += QB) {
        const int u = idx / half, i = idx - u * half;
        const int t = k0 + u;
        if (t < T) {
          const float c = rope_cos[(int64_t)t * half + i], s = rope_sin[(int64_t)t * half + i];
          const float a = Ks[u * DP + i], bb = Ks[u * DP + i + half];
          Ks[u * DP + i] = a * c - bb * s;
          Ks[u * DP + i + half] = bb * c + a * s;
        }
      }
      __syncthreads();
    }
    const int nk = min(KT, T - k0);
    float sc[KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int u = 0; u < KT; ++u) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(&Ks[u * DP + j]);
        a = fmaf(qr[j], kk.x, a);
        a = fmaf(qr[j + 1], kk.y, a);
        a = fmaf(qr[j + 2], kk.z, a);
        a = fmaf(qr[j + 3], kk.w, a);
      }
      if (lengths && !(q_valid && (k0 + u) < len)) a += -1e4f;
      if (u >= nk) a = -INFINITY;
      sc[u] = a;
      tmax = fmaxf(tmax, a);
    }
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);  // m = -inf on the first tile -> 0
    l *= corr;
#pragma unroll
    for (int j = 0; j < D; ++j) acc[j] *= corr;
#pragma unroll
    for (int u = 0; u < KT; ++u) {
      float pu = expf(sc[u] - m_new);  // -inf -> 0 for padded keys
      l += pu;
      // dropout acts on the normalised probabilities: the denominator keeps every key (SDPA dropout_p)
      if (ds.seed) pu = drop_keep(seed, ds.site, drow + (unsigned long long)(k0 + u), ds.thresh) ? pu * ds.inv_keep : 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[u * DP + j]);
        acc[j] = fmaf(pu, vv.x, acc[j]);
        acc[j + 1] = fmaf(pu, vv.y, acc[j + 1]);
        acc[j + 2] = fmaf(pu, vv.z, acc[j + 2]);
        acc[j + 3] = fmaf(pu, vv.w, acc[j + 3]);
      }
    }
    m = m_new;
    __syncthreads();
  }
  if (q_ok) {
    const float inv = 1.0f / l;
    float* __restrict__ ob = o + (int64_t)b * o_bs + (int64_t)h * D * T;
#pragma unroll
    for (int j = 0; j < D; ++j) ob[(int64_t)j * T + tq] = acc[j] * inv;
    if (lse) lse[((int64_t)b * gridDim.y + h) * T + tq] = m + logf(l);
  }
}

__global__ void rope_table_kernel(float* __restrict__ c, float* __restrict__ s, int T, int half,
                                  int d_rot, float base) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * half) return;
  const int t = idx / half, i = idx - t * half;
  // theta_i = 1 / base^(2i/d_rot) evaluated in fp64, rounded to fp32 like the
  // reference's fp32 table; the angle t*theta is an fp32 product (text_encoder.py:118-126)
  const float theta = (float)(1.0 / pow((double)base, (double)(2 * i) / (double)d_rot));
  const float ang = (float)t * theta;
  c[idx] = (float)cos((double)ang);
  s[idx] = (float)sin((double)ang);
}

}  // namespace sty

using namespace sty;

extern "C" int sty_rope_table(float* cos_out, float* sin_out, int T, int d_rot, float base,
                              sty_stream_t stream) {
  STY_REQUIRE(cos_out && sin_out && T > 0 && d_rot >= 2 && d_rot % 2 == 0, "rope_table: bad argument");
  const int half = d_rot / 2;
  rope_table_kernel<<<cdiv((int64_t)T * half, 256), 256, 0, as_stream(stream)>>>(cos_out, sin_out, T,
                                                                                half, d_rot, base);
  STY_CHECK_LAUNCH("rope_table");
  return STY_OK;
}

namespace sty {
int attention_umma_launch(const float* q, const float* k, const float* v, int64_t qkv_bs, float* o, int64_t o_bs,
                          int B, int H, int T, float scale, float* lse, cudaStream_t st);  // attention_umma.cu
}

static int attention_launch(const float* q, const float* k, const float* v, int64_t qkv_bs,
                            float* o, int64_t o_bs, const int64_t* lengths,
                            const float* rope_cos, const float* rope_sin, int d_rot, int B,
                            int H, int D, int T, float scale, float* lse, sty_stream_t stream,
                            const sty_dropout* drop = nullptr) {
  STY_REQUIRE(q && k && v && o, "attention: null pointer");
  STY_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "attention: dropout p must be in [0,1)");
  const DropSpec ds = make_drop(drop);
  STY_REQUIRE(B > 0 && H > 0 && T > 0, "attention: bad shape");
  STY_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "attention: need both rope tables");
  STY_REQUIRE(!rope_cos || (d_rot >= 2 && d_rot % 2 == 0 && d_rot <= D), "attention: bad d_rot=%d", d_rot);
  STY_REQUIRE(H <= 65535 && B <= 65535, "attention: grid too large");
  cudaStream_t st = as_stream(stream);
  if (D == 16) {
    constexpr int QB = 64;
    dim3 grid(cdiv(T, QB), H, B);
    if (rope_cos) {
      STY_REQUIRE(d_rot == 8, "attention: D=16 is built with d_rot=8 (got %d)", d_rot);
      attention_kernel<16, 32, QB, 4><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, o_bs, lengths,
                                                           rope_cos, rope_sin, T, scale, lse, ds);
    } else {
      attention_kernel<16, 32, QB, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, o_bs, lengths,
                                                           rope_cos, rope_sin, T, scale, lse, ds);
    }
  } else if (D == 64 && !rope_cos && !lengths && !ds.seed && T >= 64 && !getenv("STYLISH_B200_ATTN_SIMT")) {
    return attention_umma_launch(q, k, v, qkv_bs, o, o_bs, B, H, T, scale, lse, st);  // tcgen05 path
  } else if (D == 64) {
    constexpr int QB = 128;
    dim3 grid(cdiv(T, QB), H, B);
    STY_REQUIRE(!rope_cos, "attention: D=64 is built without RoPE");
    attention_kernel<64, 16, QB, 0><<<grid, QB, 0, st>>>(q, k, v, qkv_bs, o, o_bs, lengths, rope_cos,
                                                         rope_sin, T, scale, lse, ds);
  } else {
    set_error("attention: unsupported head dim %d (built: 16, 64)", D);
    return STY_ERR_BAD_ARG;
  }
  STY_CHECK_LAUNCH("attention");
  return STY_OK;
}

extern "C" int sty_attention_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs,
                                 float* o, int64_t o_bs, const int64_t* lengths,
                                 const float* rope_cos, const float* rope_sin, int d_rot, int B,
                                 int H, int D, int T, float scale, sty_stream_t stream) {
  return attention_launch(q, k, v, qkv_bs, o, o_bs, lengths, rope_cos, rope_sin, d_rot, B, H, D, T, scale,
                          nullptr, stream);
}

extern "C" int sty_attention_lse_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs,
                                     float* o, int64_t o_bs, const int64_t* lengths,
                                     const float* rope_cos, const float* rope_sin, int d_rot, int B,
                                     int H, int D, int T, float scale, float* lse, sty_stream_t stream) {
  STY_REQUIRE(lse, "attention_lse: null lse");
  return attention_launch(q, k, v, qkv_bs, o, o_bs, lengths, rope_cos, rope_sin, d_rot, B, H, D, T, scale, lse,
                          stream);
}

extern "C" int sty_attention_drop_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs, float* o,
                                      int64_t o_bs, const int64_t* lengths, const float* rope_cos,
                                      const float* rope_sin, int d_rot, int B, int H, int D, int T, float scale,
                                      float* lse, const sty_dropout* drop, sty_stream_t stream) {
  STY_REQUIRE(lse, "attention_drop: null lse");
  return attention_launch(q, k, v, qkv_bs, o, o_bs, lengths, rope_cos, rope_sin, d_rot, B, H, D, T, scale, lse,
                          stream, drop);
}
