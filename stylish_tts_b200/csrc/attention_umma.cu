// Multi-head attention core on the tensor cores (tcgen05.mma, accumulators in TMEM) for 64-wide heads
// without mask / RoPE — the conformer attention of the vocoder (conformer.py:112-131: 8 heads x 64,
// T ~ 800 frames).  Flash-style: one CTA = 128 queries of one (batch, head); per 64-key tile
//   S = (scale*Q) K^T  (M=128, N=64, K=64)   -> TMEM
//   online softmax on the S rows (one thread per query row, tcgen05.ld)
//   P (bf16 hi|lo) -> shared memory, O_tile = P V  (M=128, N=64, K=64) -> TMEM -> rescaled register accumulator
// fp32 operands are split into bf16 (hi, lo) and three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32
// ("bf16x3"), like the convolution kernels.  Q, K are staged K-major ([d/8][token][8]); V is staged
// [d/8][key][8] and used as an MN-major B operand (N = d, K = key), so all three tensors use the same
// coalesced 8-channel gather.  Two CTAs fit one SM (96 KB shared memory, 128 TMEM columns each).
#include <math.h>

#include "umma.cuh"

namespace sty {
namespace {

constexpr int kAD = 64;    // head dim
constexpr int kAQ = 128;   // queries per CTA
constexpr int kAK = 64;    // keys per tile
constexpr int kAThreads = 128;

// stage rows [t0, t0+rows) x 64 channels of one head as bf16 hi | lo planes, dst[(split*8 + d8)*rows + row]
__device__ __forceinline__ void stage_head(uint4* __restrict__ dst, const float* __restrict__ src, int T, int t0,
                                           int rows, float mul, int tid) {
  const int n_items = 8 * rows;  // (d8, row)
  for (int i0 = tid; i0 < n_items; i0 += kAThreads * 4) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = i0 + u * kAThreads;
      const int d8 = item / rows, row = item - d8 * rows;
      const int t = t0 + row;
      const bool ok = item < n_items && t < T;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[u][j] = ok ? src[(int64_t)(d8 * 8 + j) * T + t] * mul : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int item = i0 + u * kAThreads;
      if (item >= n_items) continue;
      const int d8 = item / rows, row = item - d8 * rows;
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = v[u][2 * j], a1 = v[u][2 * j + 1];
        h[j] = pack_bf16(a0, a1);
        l[j] = pack_bf16(a0 - __uint_as_float(h[j] << 16), a1 - __uint_as_float(h[j] & 0xffff0000u));
      }
      dst[(0 * 8 + d8) * rows + row] = make_uint4(h[0], h[1], h[2], h[3]);
      dst[(1 * 8 + d8) * rows + row] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// token-major source: row t holds the 64 features of this head contiguously, rows `ld` floats apart
__device__ __forceinline__ void stage_head_tm(uint4* __restrict__ dst, const float* __restrict__ src, int64_t ld, int T,
                                              int t0, int rows, float mul, int tid) {
  const int n_items = 8 * rows;  // (row, d8): d8 fastest so that a warp reads whole 256-byte rows
  for (int item = tid; item < n_items; item += kAThreads) {
    const int row = item >> 3, d8 = item & 7;
    const int t = t0 + row;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (t < T) {
      const float* sp = src + (int64_t)t * ld + d8 * 8;
      a = *reinterpret_cast<const float4*>(sp);
      b = *reinterpret_cast<const float4*>(sp + 4);
    }
    const float v[8] = {a.x * mul, a.y * mul, a.z * mul, a.w * mul, b.x * mul, b.y * mul, b.z * mul, b.w * mul};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
      l[j] = pack_bf16(v[2 * j] - __uint_as_float(h[j] << 16), v[2 * j + 1] - __uint_as_float(h[j] & 0xffff0000u));
    }
    dst[(0 * 8 + d8) * rows + row] = make_uint4(h[0], h[1], h[2], h[3]);
    dst[(1 * 8 + d8) * rows + row] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// TM = false: q, k, v channel-major (B, H*64, T) fp32, output channel-major fp32 (the conformer).
// TM = true : q, k, v token-major rows of `qkv_bs` floats ([B*T, ld], head h at columns h*64 of each pointer),
//             output = bf16 hi | lo planes [2][o_bs rows][H*64] (the next GEMM's operand; o_bs = padded row count).
template <bool TM>
__global__ void __launch_bounds__(kAThreads, 2)
attention_umma_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      int64_t qkv_bs, float* __restrict__ o, int64_t o_bs, int T, float scale,
                      float* __restrict__ lse) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* Qs = reinterpret_cast<uint4*>(smem_raw);  // [2][8][128]
  uint4* Ks = Qs + 2 * 8 * kAQ;                    // [2][8][64]
  uint4* Vs = Ks + 2 * 8 * kAK;                    // [2][8][64]
  uint4* Ps = Vs + 2 * 8 * kAK;                    // [2][8 key groups][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 2 * (kAK / 8) * kAQ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kAQ;
  const int64_t hoff = TM ? (int64_t)b * T * qkv_bs + (int64_t)h * kAD : (int64_t)b * qkv_bs + (int64_t)h * kAD * T;
  const float* __restrict__ qb = q + hoff;
  const float* __restrict__ kb = k + hoff;
  const float* __restrict__ vb = v + hoff;

  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (TM) stage_head_tm(Qs, qb, qkv_bs, T, q0, kAQ, scale, tid);
  else stage_head(Qs, qb, T, q0, kAQ, scale, tid);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot, tmem_o = tmem_s + 64;
  // S = Q K^T : A, B K-major, M = 128, N = 64;   O = P V : A K-major, B MN-major, M = 128, N = 64
  const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kAK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_o = idesc_s | (1u << 16);

  float m_run = -INFINITY, l_run = 0.f;
  float acc[kAD];
#pragma unroll
  for (int j = 0; j < kAD; ++j) acc[j] = 0.f;
  const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);

  const int n_tiles = (T + kAK - 1) / kAK;
  for (int kt = 0; kt < n_tiles; ++kt) {
    const int k0 = kt * kAK;
    // the previous tile's MMAs have completed (waited below), so K/V/P buffers are free
    if (TM) {
      stage_head_tm(Ks, kb, qkv_bs, T, k0, kAK, 1.f, tid);
      stage_head_tm(Vs, vb, qkv_bs, T, k0, kAK, 1.f, tid);
    } else {
      stage_head(Ks, kb, T, k0, kAK, 1.f, tid);
      stage_head(Vs, vb, T, k0, kAK, 1.f, tid);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t ad = make_desc(smem_u32(Qs), (uint32_t)kAQ, 8u);
      const uint64_t bd = make_desc(smem_u32(Ks), (uint32_t)kAK, 8u);
      const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
      const uint32_t a_lo = (uint32_t)ad, b_lo = (uint32_t)bd;
#pragma unroll
      for (uint32_t ks = 0; ks < kAD / 16; ++ks) {
        const uint32_t ak = a_lo + ks * 2u * kAQ, bk = b_lo + ks * 2u * kAK;
        umma_bf16_w(tmem_s, ak, a_hi, bk, b_hi, idesc_s, ks == 0 ? 0u : 1u);
        umma_bf16_w(tmem_s, ak + 8u * kAQ, a_hi, bk, b_hi, idesc_s, 1u);   // lo * hi
        umma_bf16_w(tmem_s, ak, a_hi, bk + 8u * kAK, b_hi, idesc_s, 1u);   // hi * lo
      }
      umma_commit(&bars[0]);
    }
    mbar_wait(&bars[0], (uint32_t)(kt & 1));
    tc_fence_after();
    // ---- online softmax on this thread's row of S
    float s[kAK];
    tmem_ld32(tmem_s + lane_addr, s);
    tmem_ld32(tmem_s + lane_addr + 32u, s + 32);
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < kAK; ++j) {
      if (k0 + j >= T) s[j] = -INFINITY;
      tmax = fmaxf(tmax, s[j]);
    }
    const float m_new = fmaxf(m_run, tmax);
    const float corr = __expf(m_run - m_new);  // exp(-inf) = 0 on the first tile
    float psum = 0.f;
#pragma unroll
    for (int j = 0; j < kAK; ++j) {
      s[j] = __expf(s[j] - m_new);
      psum += s[j];
    }
    l_run = fmaf(l_run, corr, psum);
    m_run = m_new;
    // P as bf16 hi | lo, K-major A operand: Ps[(split*8 + key/8)*128 + row]
#pragma unroll
    for (int g = 0; g < kAK / 8; ++g) {
      uint32_t hh[4], ll[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = s[g * 8 + 2 * j], a1 = s[g * 8 + 2 * j + 1];
        hh[j] = pack_bf16(a0, a1);
        ll[j] = pack_bf16(a0 - __uint_as_float(hh[j] << 16), a1 - __uint_as_float(hh[j] & 0xffff0000u));
      }
      Ps[(0 * 8 + g) * kAQ + tid] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
      Ps[(1 * 8 + g) * kAQ + tid] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t ad = make_desc(smem_u32(Ps), (uint32_t)kAQ, 8u);          // K-major: k-groups 128 rows apart
      const uint64_t bd = make_desc(smem_u32(Vs), 8u, (uint32_t)kAK);           // MN-major: lbo = 8 keys, sbo = d group
      const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
      const uint32_t a_lo = (uint32_t)ad, b_lo = (uint32_t)bd;
#pragma unroll
      for (uint32_t ks = 0; ks < kAK / 16; ++ks) {
        const uint32_t ak = a_lo + ks * 2u * kAQ, bk = b_lo + ks * 16u;
        umma_bf16_w(tmem_o, ak, a_hi, bk, b_hi, idesc_o, ks == 0 ? 0u : 1u);
        umma_bf16_w(tmem_o, ak + 8u * kAQ, a_hi, bk, b_hi, idesc_o, 1u);  // P lo * V hi
        umma_bf16_w(tmem_o, ak, a_hi, bk + 8u * kAK, b_hi, idesc_o, 1u);  // P hi * V lo
      }
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], (uint32_t)(kt & 1));
    tc_fence_after();
    float pv[32];
    tmem_ld32(tmem_o + lane_addr, pv);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], corr, pv[j]);
    tmem_ld32(tmem_o + lane_addr + 32u, pv);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[32 + j] = fmaf(acc[32 + j], corr, pv[j]);
    tc_fence_before();  // TMEM reads done before the next tile's MMAs overwrite S / O
  }
  const int tq = q0 + tid;
  if (tq < T) {
    const float inv = 1.0f / l_run;
    if constexpr (TM) {
      const int C = gridDim.y * kAD;
      const int64_t row = (int64_t)b * T + tq;
      __nv_bfloat16* os = reinterpret_cast<__nv_bfloat16*>(o);
      uint4* oh = reinterpret_cast<uint4*>(os + row * C + h * kAD);
      uint4* ol = reinterpret_cast<uint4*>(os + (o_bs + row) * C + h * kAD);
#pragma unroll
      for (int g = 0; g < kAD / 8; ++g) {
        uint32_t hh[4], ll[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = acc[g * 8 + 2 * e] * inv, v1 = acc[g * 8 + 2 * e + 1] * inv;
          hh[e] = pack_bf16(v0, v1);
          ll[e] = pack_bf16(v0 - __uint_as_float(hh[e] << 16), v1 - __uint_as_float(hh[e] & 0xffff0000u));
        }
        oh[g] = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        ol[g] = make_uint4(ll[0], ll[1], ll[2], ll[3]);
      }
    } else {
      float* __restrict__ ob = o + (int64_t)b * o_bs + (int64_t)h * kAD * T;  // o_bs may differ from qkv_bs
#pragma unroll
      for (int j = 0; j < kAD; ++j) ob[(int64_t)j * T + tq] = acc[j] * inv;
      if (lse) lse[((int64_t)b * gridDim.y + h) * T + tq] = m_run + logf(l_run);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_s, 128);
}

}  // namespace

int attention_umma_launch(const float* q, const float* k, const float* v, int64_t qkv_bs, float* o, int64_t o_bs,
                          int B, int H, int T, float scale, float* lse, cudaStream_t st) {
  const size_t smem = (size_t)(2 * 8 * kAQ + 2 * 2 * 8 * kAK + 2 * (kAK / 8) * kAQ) * 16 + 64;
  cudaFuncSetAttribute(attention_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(T, kAQ), H, B);
  attention_umma_kernel<false><<<grid, kAThreads, smem, st>>>(q, k, v, qkv_bs, o, o_bs, T, scale, lse);
  STY_CHECK_LAUNCH("attention_umma");
  return STY_OK;
}

}  // namespace sty

// token-major 64-wide attention of the diffusion denoiser (no mask): qkv rows [B*T, ld] fp32 with q | k | v at
// column offsets 0 / H*64 / 2*H*64; output bf16 hi | lo planes [2][M_pad][H*64]
extern "C" int sty_attention_tokens_fwd(const float* qkv, int64_t ld, void* out_split, int64_t M_pad, int B, int H,
                                        int T, float scale, sty_stream_t stream) {
  using namespace sty;
  STY_REQUIRE(qkv && out_split && B > 0 && H > 0 && T > 0 && ld >= 3 * H * kAD && (ld & 3) == 0 &&
                  M_pad >= (int64_t)B * T, "attention_tokens: bad argument");
  const size_t smem = (size_t)(2 * 8 * kAQ + 2 * 2 * 8 * kAK + 2 * (kAK / 8) * kAQ) * 16 + 64;
  cudaFuncSetAttribute(attention_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(T, kAQ), H, B);
  attention_umma_kernel<true><<<grid, kAThreads, smem, as_stream(stream)>>>(
      qkv, qkv + H * kAD, qkv + 2 * H * kAD, ld, reinterpret_cast<float*>(out_split), M_pad, T, scale, nullptr);
  STY_CHECK_LAUNCH("attention_tokens");
  return STY_OK;
}

namespace sty {

}  // namespace sty
