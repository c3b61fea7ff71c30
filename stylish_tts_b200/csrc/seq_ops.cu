// Sequence-level ops of the duration / pitch-energy predictors: attention for arbitrary
// head sizes, the duration head, soft durations and the soft alignment matrix.
#include <math.h>

#include "common.cuh"

namespace sty {

// ------------------------------------------------------- attention, any head size
// CTA = (b, h, 32 queries).  Q tile, and K/V tiles of 32 keys, live in shared memory
// (row pitch D+1).  Thread (qi, sl): qi = tid % 32 owns query row qi; sl = tid / 32 (8
// slices) owns keys {sl, sl+8, ..} of a tile for the score pass and the output dims
// {sl, sl+8, ...} for the P.V pass.  Streaming softmax state per row in shared memory.
template <int DMAX>
__global__ void __launch_bounds__(256)
attention_generic_kernel(const float* __restrict__ q, int64_t q_bs, const float* __restrict__ k,
                         const float* __restrict__ v, int64_t kv_bs, float* __restrict__ o,
                         int64_t o_bs, const int64_t* __restrict__ lengths,
                         const float* __restrict__ rope_cos, const float* __restrict__ rope_sin,
                         int d_rot, int D, int T, float scale) {
  constexpr int QT = 32, KT = 32, NSL = 8, ND = DMAX / NSL;
  extern __shared__ float sm[];
  const int DP = D + 1;
  float* Qs = sm;                  // [QT][DP]
  float* Ks = Qs + QT * DP;        // [KT][DP]
  float* Vs = Ks + KT * DP;        // [KT][DP]
  float* Ps = Vs + KT * DP;        // [QT][KT+1]
  float* row_m = Ps + QT * (KT + 1);  // [QT]
  float* row_l = row_m + QT;          // [QT]
  float* row_c = row_l + QT;          // [QT] correction factor of the current tile
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int tid = threadIdx.x, qi = tid & 31, sl = tid >> 5;
  const int half = d_rot >> 1;
  const int len = lengths ? (int)lengths[b] : T;
  const float* __restrict__ qb = q + (int64_t)b * q_bs + (int64_t)h * D * T;
  const float* __restrict__ kb = k + (int64_t)b * kv_bs + (int64_t)h * D * T;
  const float* __restrict__ vb = v + (int64_t)b * kv_bs + (int64_t)h * D * T;

  auto load_tile = [&](float* dst, const float* src, int t0, int n_rows, bool rope, float mul) {
    for (int idx = tid; idx < n_rows * D; idx += 256) {
      const int j = idx / n_rows, r = idx - j * n_rows;  // consecutive threads: consecutive time
      const int t = t0 + r;
      dst[r * DP + j] = (t < T) ? src[(int64_t)j * T + t] * mul : 0.f;
    }
    __syncthreads();
    if (rope) {
      for (int idx = tid; idx < n_rows * half; idx += 256) {
        const int r = idx / half, i = idx - r * half;
        const int t = t0 + r;
        if (t < T) {
          const float c = rope_cos[(int64_t)t * half + i], s = rope_sin[(int64_t)t * half + i];
          const float a = dst[r * DP + i], bb = dst[r * DP + i + half];
          dst[r * DP + i] = a * c - bb * s;
          dst[r * DP + i + half] = bb * c + a * s;
        }
      }
      __syncthreads();
    }
  };

  load_tile(Qs, qb, q0, QT, rope_cos != nullptr, 1.f);
  if (tid < QT) {
    row_m[tid] = -INFINITY;
    row_l[tid] = 0.f;
  }
  float acc[ND];
#pragma unroll
  for (int i = 0; i < ND; ++i) acc[i] = 0.f;
  const bool q_valid = (q0 + qi) < len;
  __syncthreads();

  for (int k0 = 0; k0 < T; k0 += KT) {
    load_tile(Ks, kb, k0, KT, rope_cos != nullptr, 1.f);
    load_tile(Vs, vb, k0, KT, false, 1.f);
    // scores for keys sl, sl+8, sl+16, sl+24 of this tile
#pragma unroll
    for (int u = 0; u < KT / NSL; ++u) {
      const int kk = sl + u * NSL;
      float a = 0.f;
      for (int j = 0; j < D; ++j) a = fmaf(Qs[qi * DP + j], Ks[kk * DP + j], a);
      a *= scale;
      if (lengths && !(q_valid && (k0 + kk) < len)) a += -1e4f;
      if (k0 + kk >= T) a = -INFINITY;
      Ps[qi * (KT + 1) + kk] = a;
    }
    __syncthreads();
    if (tid < QT) {  // one thread per query row: new max, correction, probabilities, row sum
      float mx = row_m[tid];
      const float m_old = mx;
      for (int kk = 0; kk < KT; ++kk) mx = fmaxf(mx, Ps[tid * (KT + 1) + kk]);
      float sum = 0.f;
      for (int kk = 0; kk < KT; ++kk) {
        const float pv = expf(Ps[tid * (KT + 1) + kk] - mx);
        Ps[tid * (KT + 1) + kk] = pv;
        sum += pv;
      }
      const float corr = expf(m_old - mx);
      row_c[tid] = corr;
      row_l[tid] = row_l[tid] * corr + sum;
      row_m[tid] = mx;
    }
    __syncthreads();
    {
      const float corr = row_c[qi];
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        const int j = sl + i * NSL;
        if (j < D) {
          float a = acc[i] * corr;
          for (int kk = 0; kk < KT; ++kk) a = fmaf(Ps[qi * (KT + 1) + kk], Vs[kk * DP + j], a);
          acc[i] = a;
        }
      }
    }
    __syncthreads();
  }
  const int tq = q0 + qi;
  if (tq < T) {
    const float inv = 1.0f / row_l[qi];
    float* __restrict__ ob = o + (int64_t)b * o_bs + (int64_t)h * D * T;
#pragma unroll
    for (int i = 0; i < ND; ++i) {
      const int j = sl + i * NSL;
      if (j < D) ob[(int64_t)j * T + tq] = acc[i] * inv;
    }
  }
}

// ----------------------------------------------------------------- duration head
__global__ void duration_head_kernel(const float* __restrict__ x, const int64_t* __restrict__ lengths,
                                     float* __restrict__ out, int NC, int T) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float m = (t < lengths[b]) ? 1.f : 0.f;
  float cum = 0.f;
  for (int c = 0; c < NC; ++c) {
    float v = x[((int64_t)b * NC + c) * T + t];
    if (c > 0) v = fabsf(v);
    cum += v;
    out[((int64_t)b * T + t) * NC + c] = -fabsf(cum) * m;
  }
}

// --------------------------------------------------------------- soft durations
// one CTA per batch element; also reduces the rounded total frame count (max over b)
__global__ void __launch_bounds__(256)
soft_duration_kernel(const float* __restrict__ pred, const int64_t* __restrict__ lengths,
                     const float* __restrict__ table, float* __restrict__ dur,
                     int32_t* __restrict__ total, int T, int NC) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int len = (int)lengths[b];
  float part = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float* __restrict__ p = pred + ((int64_t)b * T + t) * NC;
    float mx = -INFINITY;
    for (int c = 0; c < NC; ++c) mx = fmaxf(mx, p[c]);
    float se = 0.f;
    for (int c = 0; c < NC; ++c) se += expf(p[c] - mx);
    float num = 0.f, den = 0.f;
    for (int c = 0; c < NC; ++c) {
      const float pr = expf(p[c] - mx) / se;
      num = fmaf(pr, table[c], num);
      den += pr;
    }
    const float d = (t < len) ? num / (den + 1e-9f) : 0.f;
    dur[(int64_t)b * T + t] = d;
    part += d;
  }
  const float sum = block_sum(part, red);
  if (threadIdx.x == 0) atomicMax(total, (int32_t)rintf(sum));
}

// -------------------------------------------------------------------- alignment
__global__ void __launch_bounds__(128)
alignment_kernel(const float* __restrict__ duration, float* __restrict__ alignment, int T, int F) {
  extern __shared__ float sm[];
  float* dur = sm;          // [T]
  float* lower = sm + T;    // [T]  (already -3)
  float* upper = sm + 2 * T;  // [T]  (already +3)
  float* mean = sm + 3 * T;
  const int b = blockIdx.y;
  for (int t = threadIdx.x; t < T; t += blockDim.x) dur[t] = duration[(int64_t)b * T + t];
  __syncthreads();
  if (threadIdx.x == 0) {  // sequential fp32 cumsum, the order torch.cumsum uses on the CPU
    float cum = 0.f;
    for (int t = 0; t < T; ++t) {
      cum += dur[t];
      const float lo = cum - dur[t];
      mean[t] = (lo + cum) / 2.f;
      lower[t] = lo - 3.f;
      upper[t] = cum + 3.f;
    }
  }
  __syncthreads();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float ff = (float)f;
  auto val = [&](int t) {
    const float x = ff - mean[t];
    const float r = x * 2.f / (dur[t] + 6.f);
    float a = 1.f - r * r;
    a = (ff > lower[t] && ff < upper[t]) ? a : 0.f;
    return fmaxf(a, 0.f);
  };
  float mx = -INFINITY;
  for (int t = 0; t < T; ++t) mx = fmaxf(mx, val(t));
  float se = 0.f;
  for (int t = 0; t < T; ++t) se += expf(val(t) - mx);
  const float inv = 1.0f / se;
  float* __restrict__ ab = alignment + (int64_t)b * T * F + f;
  for (int t = 0; t < T; ++t) ab[(int64_t)t * F] = expf(val(t) - mx) * inv;
}

}  // namespace sty

using namespace sty;

extern "C" int sty_attention_generic_fwd(const float* q, int64_t q_bs, const float* k, const float* v,
                                         int64_t kv_bs, float* o, int64_t o_bs, const int64_t* lengths,
                                         const float* rope_cos, const float* rope_sin, int d_rot,
                                         int B, int H, int D, int T, float scale, sty_stream_t stream) {
  STY_REQUIRE(q && k && v && o, "attention_generic: null pointer");
  STY_REQUIRE(B > 0 && H > 0 && T > 0 && D > 0 && D <= 256, "attention_generic: bad shape (D<=256)");
  STY_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "attention_generic: need both rope tables");
  STY_REQUIRE(!rope_cos || (d_rot >= 2 && d_rot % 2 == 0 && d_rot <= D), "attention_generic: bad d_rot");
  const size_t smem = ((size_t)(32 + 64) * (D + 1) + 32 * 33 + 96) * sizeof(float);
  auto kern = attention_generic_kernel<256>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(T, 32), H, B);
  kern<<<grid, 256, smem, as_stream(stream)>>>(q, q_bs, k, v, kv_bs, o, o_bs, lengths, rope_cos, rope_sin,
                                               d_rot, D, T, scale);
  STY_CHECK_LAUNCH("attention_generic");
  return STY_OK;
}

extern "C" int sty_duration_head_fwd(const float* x, const int64_t* lengths, float* out, int B, int NC,
                                     int T, sty_stream_t stream) {
  STY_REQUIRE(x && lengths && out && B > 0 && NC > 0 && T > 0, "duration_head: bad argument");
  dim3 grid(cdiv(T, 128), B);
  duration_head_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, lengths, out, NC, T);
  STY_CHECK_LAUNCH("duration_head");
  return STY_OK;
}

extern "C" int sty_soft_duration_fwd(const float* pred, const int64_t* lengths, const float* table,
                                     float* dur, int32_t* total, int B, int T, int NC,
                                     sty_stream_t stream) {
  STY_REQUIRE(pred && lengths && table && dur && total && B > 0 && T > 0 && NC > 0,
              "soft_duration: bad argument");
  cudaMemsetAsync(total, 0, sizeof(int32_t), as_stream(stream));
  soft_duration_kernel<<<B, 256, 0, as_stream(stream)>>>(pred, lengths, table, dur, total, T, NC);
  STY_CHECK_LAUNCH("soft_duration");
  return STY_OK;
}

extern "C" int sty_alignment_fwd(const float* duration, float* alignment, int B, int T, int F,
                                 sty_stream_t stream) {
  STY_REQUIRE(duration && alignment && B > 0 && T > 0 && F > 0, "alignment: bad argument");
  STY_REQUIRE(T <= 2048, "alignment: T too large for the staging buffer");
  dim3 grid(cdiv(F, 128), B);
  alignment_kernel<<<grid, 128, 4 * T * sizeof(float), as_stream(stream)>>>(duration, alignment, T, F);
  STY_CHECK_LAUNCH("alignment");
  return STY_OK;
}
