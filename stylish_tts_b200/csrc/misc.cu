// Small glue kernels: embedding gather, batched matmul, GLU.
#include "common.cuh"

namespace sty {

// out[b,c,t] = emb[tok[b,t], c] * scale.  A 32x32 tile is transposed through
// shared memory so both the gather (rows of C floats) and the (B,C,T) store are
// coalesced.
__global__ void __launch_bounds__(256)
embed_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ lengths,
             const float* __restrict__ emb, float* __restrict__ out, int T, int C, int n_tokens,
             float scale) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, c = c0 + tx;
    float v = 0.f;
    if (t < T && c < C && (!lengths || t < lengths[b])) {
      int64_t tok = tokens[(int64_t)b * T + t];
      tok = tok < 0 ? 0 : (tok >= n_tokens ? n_tokens - 1 : tok);
      v = emb[tok * C + c] * scale;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, t = t0 + tx;
    if (t < T && c < C) out[((int64_t)b * C + c) * T + t] = tile[tx][r];
  }
}

// C[b] = A[b] @ Bm[b]; 64x64 tile, BK = 16, 256 threads, 4x4 register tile.
__global__ void __launch_bounds__(256)
bmm_kernel(const float* __restrict__ A, int64_t a_bs, const float* __restrict__ Bm, int64_t b_bs,
           float* __restrict__ C, int64_t c_bs, int M, int N, int K) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;  // 16 x 16 threads
  const float* __restrict__ Ab = A + (int64_t)b * a_bs;
  const float* __restrict__ Bb = Bm + (int64_t)b * b_bs;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int idx = tid; idx < BM * BK; idx += 256) {
      const int kk = idx % BK, mm = idx / BK;  // consecutive threads walk K (row-major A)
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? Ab[(int64_t)m * K + k] : 0.f;
    }
    for (int idx = tid; idx < BK * BN; idx += 256) {
      const int nn = idx % BN, kk = idx / BN;
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? Bb[(int64_t)k * N + n] : 0.f;
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
  float* __restrict__ Cb = C + (int64_t)b * c_bs;
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

__global__ void __launch_bounds__(256)
glu_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int T) {
  const int64_t n = (int64_t)C * T;
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = x[(int64_t)b * 2 * n + i];
  const float g = x[(int64_t)b * 2 * n + n + i];
  y[(int64_t)b * n + i] = a * (1.0f / (1.0f + expf(-g)));
}

__global__ void sequence_mask_kernel(const int64_t* __restrict__ lengths, float* __restrict__ out,
                                     int B, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T) return;
  const int b = (int)(i / T), t = (int)(i - (int64_t)b * T);
  out[i] = (t < lengths[b]) ? 1.f : 0.f;
}

// out[b,c,t] = x[b,c,t] * mask[b,t] * scale
__global__ void scale_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                  float* __restrict__ out, int C, int T, float scale) {
  const int b = blockIdx.y;
  const int64_t n = (int64_t)C * T;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = (int)(i % T);
  out[(int64_t)b * n + i] = x[(int64_t)b * n + i] * mask[(int64_t)b * T + t] * scale;
}

// out[b,s,t] = v[b,s]  (style vector repeated over time; prosody_encoder.py:67-68)
__global__ void broadcast_rows_kernel(const float* __restrict__ v, float* __restrict__ out, int64_t out_bs,
                                      int S, int T) {
  const int b = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)S * T) return;
  out[(int64_t)b * out_bs + i] = v[(int64_t)b * S + (int)(i / T)];
}

}  // namespace sty

using namespace sty;

extern "C" int sty_scale_mask_fwd(const float* x, const float* mask, float* out, int B, int C, int T,
                                  float scale, sty_stream_t stream) {
  STY_REQUIRE(x && mask && out && B > 0 && C > 0 && T > 0, "scale_mask: bad argument");
  dim3 grid(cdiv((int64_t)C * T, 256), B);
  scale_mask_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, mask, out, C, T, scale);
  STY_CHECK_LAUNCH("scale_mask");
  return STY_OK;
}

extern "C" int sty_broadcast_rows_fwd(const float* v, float* out, int64_t out_bs, int B, int S, int T,
                                      sty_stream_t stream) {
  STY_REQUIRE(v && out && B > 0 && S > 0 && T > 0, "broadcast_rows: bad argument");
  dim3 grid(cdiv((int64_t)S * T, 256), B);
  broadcast_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(v, out, out_bs, S, T);
  STY_CHECK_LAUNCH("broadcast_rows");
  return STY_OK;
}

extern "C" int sty_sequence_mask_fwd(const int64_t* lengths, float* out, int B, int T,
                                     sty_stream_t stream) {
  STY_REQUIRE(lengths && out && B > 0 && T > 0, "sequence_mask: bad argument");
  sequence_mask_kernel<<<cdiv((int64_t)B * T, 256), 256, 0, as_stream(stream)>>>(lengths, out, B, T);
  STY_CHECK_LAUNCH("sequence_mask");
  return STY_OK;
}

extern "C" int sty_embed_fwd(const int64_t* tokens, const int64_t* lengths, const float* emb,
                             float* out, int B, int T, int C, int n_tokens, float scale,
                             sty_stream_t stream) {
  STY_REQUIRE(tokens && emb && out && B > 0 && T > 0 && C > 0 && n_tokens > 0, "embed: bad argument");
  dim3 grid(cdiv(T, 32), cdiv(C, 32), B);
  embed_kernel<<<grid, 256, 0, as_stream(stream)>>>(tokens, lengths, emb, out, T, C, n_tokens,
                                                    scale);
  STY_CHECK_LAUNCH("embed");
  return STY_OK;
}

extern "C" int sty_bmm_fwd(const float* A, int64_t a_bs, const float* Bm, int64_t b_bs, float* C,
                           int64_t c_bs, int B, int M, int N, int K, sty_stream_t stream) {
  STY_REQUIRE(A && Bm && C && B > 0 && M > 0 && N > 0 && K > 0, "bmm: bad argument");
  dim3 grid(cdiv(N, 64), cdiv(M, 64), B);
  bmm_kernel<<<grid, 256, 0, as_stream(stream)>>>(A, a_bs, Bm, b_bs, C, c_bs, M, N, K);
  STY_CHECK_LAUNCH("bmm");
  return STY_OK;
}

extern "C" int sty_glu_fwd(const float* x, float* y, int B, int C, int T, sty_stream_t stream) {
  STY_REQUIRE(x && y && B > 0 && C > 0 && T > 0, "glu: bad argument");
  dim3 grid(cdiv((int64_t)C * T, 256), B);
  glu_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, C, T);
  STY_CHECK_LAUNCH("glu");
  return STY_OK;
}
