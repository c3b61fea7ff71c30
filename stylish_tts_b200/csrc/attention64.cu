// 64-wide multi-head attention on the tensor cores, second generation: operands pre-split once, tiles moved by the
// bulk-copy engine, a dedicated issuer warp, MMAs of the next tile queued behind the current tile's.
// (conformer attention of the vocoder, conformer.py:112-131: 8 heads x 64, T ~ 800 frames, no mask / RoPE; and the
// Transformer1d blocks of the style-diffusion denoiser.)  Forward AND backward.
//
//   prepare   q, k, v (fp32) -> bf16 hi | lo planes in the EXACT shared-memory image of one tile
//             (Q: [2 split][8 d8][128 rows][8], K / V: [2][8][64 rows][8]); q is scaled by scale*log2(e) on the way.
//             One pass over q, k, v; every 128-query CTA of the attention kernel then reads K / V without touching an
//             ALU (the first version converted each K / V tile cdiv(T,128) times, in the critical path of every tile).
//   attention one CTA = 128 queries of one (batch, head), 288 threads, 2 CTAs per SM.  Per 64-key tile
//               issuer warp : cp.async.bulk K, V tile -> shared memory (mbarrier complete_tx), one tile ahead;
//                             S = Q K^T (M=128, N=64, K=64), O_t = P V (M=128, N=64, K=64) on tcgen05, accumulators in
//                             TMEM; S of the NEXT tile is queued right behind P V of the current one
//               8 softmax warps: 2 threads per query row (TMEM lane), 32 columns each; tile maximum exchanged through
//                             shared memory + a 64-thread named barrier, partial row sums combined at the end;
//                             P = ex2(S - m) (one MUFU) as bf16 hi | lo -> shared memory; O_t folded into registers
//             bf16x3 split precision (hi*hi + lo*hi + hi*lo, fp32 accumulation) as everywhere.
//   backward  see the block comment above attn64_bwd_prepare_kernel.
#include <math.h>

#include "tma.cuh"

namespace sty {
namespace {

constexpr int kD = 64, kQ = 128, kK = 64, kThreads = 288, kIssuerWarp = 8;
constexpr int kQTileU4 = 2 * 8 * kQ, kKTileU4 = 2 * 8 * kK;  // 16-byte units per staged tile (32 KB / 16 KB)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// one MUFU.EX2 (exp2f() wraps it in a denormal-range rescue: 2 FMUL + FSETP per value)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
    l[j] = pack_bf16(v[2 * j] - __uint_as_float(h[j] << 16), v[2 * j + 1] - __uint_as_float(h[j] & 0xffff0000u));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// block = 64 rows (tokens) of one (b, h): q (half a 128-row Q tile), k, v
// TM = false: channel-major (B, H*64, T) fp32, batch stride bs;  TM = true: token-major rows of ld floats
template <bool TM>
__global__ void __launch_bounds__(256)
attn64_prepare_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      int64_t bs, uint4* __restrict__ Qw, uint4* __restrict__ Kw, uint4* __restrict__ Vw, int T,
                      int n_qb, int n_kt, float scale) {
  const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
  const int t0 = tile * 64;
  const int64_t bh = (int64_t)b * H + h;
  const int64_t off = TM ? (int64_t)b * T * bs + (int64_t)h * kD : (int64_t)b * bs + (int64_t)h * kD * T;
  uint4* qd = Qw + (bh * n_qb + (tile >> 1)) * kQTileU4 + (tile & 1) * 64;
  uint4* kd = Kw + (bh * n_kt + tile) * kKTileU4;
  uint4* vd = Vw + (bh * n_kt + tile) * kKTileU4;
  for (int item = threadIdx.x; item < 3 * 8 * 64; item += 256) {
    const int which = item / 512, r = item % 512;
    if (which != 0 && tile >= n_kt) continue;
    // TM: d8 fastest (a warp reads whole 256-byte rows); channel-major: row fastest (coalesced along t)
    const int d8 = TM ? (r & 7) : (r >> 6), row = TM ? (r >> 3) : (r & 63);
    const int t = t0 + row;
    const float* __restrict__ src = (which == 0 ? q : which == 1 ? k : v) + off;
    const float mul = which == 0 ? scale : 1.f;
    float x[8];
    if (t < T) {
      if (TM) {
        const float4 a = *reinterpret_cast<const float4*>(src + (int64_t)t * bs + d8 * 8);
        const float4 c = *reinterpret_cast<const float4*>(src + (int64_t)t * bs + d8 * 8 + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = src[(int64_t)(d8 * 8 + j) * T + t];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] *= mul;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    if (which == 0) {
      qd[(0 * 8 + d8) * kQ + row] = hi;
      qd[(1 * 8 + d8) * kQ + row] = lo;
    } else {
      uint4* d = which == 1 ? kd : vd;
      d[(0 * 8 + d8) * kK + row] = hi;
      d[(1 * 8 + d8) * kK + row] = lo;
    }
  }
}

// OUT_SPLIT = false: o fp32 channel-major (B, H*64, T), batch stride o_bs (+ optional log-sum-exp (B,H,T))
// OUT_SPLIT = true : o = bf16 hi | lo planes [2][o_bs rows][H*64], row = b*T + t (operand of the next GEMM)
// 256 threads: warp w owns TMEM lanes 32*(w&3).. (one query row per lane) and the column half w>>2 of S and of O, so a
// row's softmax is shared by two threads (they exchange the tile maximum through shared memory and a 64-thread named
// barrier; the partial sums are only combined at the end).  q arrives scaled by scale*log2(e): P = exp2(S - m).
__global__ void __launch_bounds__(kThreads, 2)
attn64_kernel_impl(const uint4* __restrict__ Qw, const uint4* __restrict__ Kw, const uint4* __restrict__ Vw,
                   float* __restrict__ o, int64_t o_bs, int T, int n_qb, int n_kt, float* __restrict__ lse,
                   const bool out_split) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* Qs = reinterpret_cast<uint4*>(smem_raw);  // [2][8][128]
  uint4* Ks = Qs + kQTileU4;                       // [2][8][64]
  uint4* Vs = Ks + kKTileU4;                       // [2][8][64]
  uint4* Ps = Vs + kKTileU4;                       // [2][8 key groups][128]
  float* pmax = reinterpret_cast<float*>(Ps + kQTileU4);  // [2 parity][2 half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(pmax + 2 * 2 * kQ);
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_o = bars + 4, *bar_p = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane;
  const int b = blockIdx.z, h = blockIdx.y, qb = blockIdx.x, H = gridDim.y;
  const int64_t bh = (int64_t)b * H + h;
  const uint4* __restrict__ kt_src = Kw + bh * n_kt * kKTileU4;
  const uint4* __restrict__ vt_src = Vw + bh * n_kt * kKTileU4;

  if (warp == 0) tmem_alloc(tmem_slot, 128);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_p, 256);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot, tmem_o = tmem_s + 64;
  // S = Q K^T : A, B K-major, M = 128, N = 64;   O = P V : A K-major, B MN-major, M = 128, N = 64
  const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_o = idesc_s | (1u << 16);

  auto issue_s = [&]() {  // one thread
    const uint64_t ad = make_desc(smem_u32(Qs), (uint32_t)kQ, 8u);
    const uint64_t bd = make_desc(smem_u32(Ks), (uint32_t)kK, 8u);
    const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
    const uint32_t a_lo = (uint32_t)ad, b_lo = (uint32_t)bd;
#pragma unroll
    for (uint32_t ks = 0; ks < kD / 16; ++ks) {
      const uint32_t ak = a_lo + ks * 2u * kQ, bk = b_lo + ks * 2u * kK;
      umma_bf16_w(tmem_s, ak, a_hi, bk, b_hi, idesc_s, ks == 0 ? 0u : 1u);
      umma_bf16_w(tmem_s, ak + 8u * kQ, a_hi, bk, b_hi, idesc_s, 1u);   // lo * hi
      umma_bf16_w(tmem_s, ak, a_hi, bk + 8u * kK, b_hi, idesc_s, 1u);   // hi * lo
    }
    umma_commit(bar_s);
  };

  auto issue_pv = [&]() {  // one thread
    const uint64_t ad = make_desc(smem_u32(Ps), (uint32_t)kQ, 8u);   // K-major: k-groups 128 rows apart
    const uint64_t bd = make_desc(smem_u32(Vs), 8u, (uint32_t)kK);    // MN-major: lbo = 8 keys, sbo = d group
    const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
    const uint32_t a_lo = (uint32_t)ad, b_lo = (uint32_t)bd;
#pragma unroll
    for (uint32_t ks = 0; ks < kK / 16; ++ks) {
      const uint32_t ak = a_lo + ks * 2u * kQ, bk = b_lo + ks * 16u;
      umma_bf16_w(tmem_o, ak, a_hi, bk, b_hi, idesc_o, ks == 0 ? 0u : 1u);
      umma_bf16_w(tmem_o, ak + 8u * kQ, a_hi, bk, b_hi, idesc_o, 1u);  // P lo * V hi
      umma_bf16_w(tmem_o, ak, a_hi, bk + 8u * kK, b_hi, idesc_o, 1u);  // P hi * V lo
    }
    umma_commit(bar_o);
  };

  float m_run = -INFINITY, l_run = 0.f;  // l_run: this thread's 32 columns only
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;

  if (warp == kIssuerWarp) {
    // ===== issuer: bulk copies and MMAs, one elected lane; the softmax warps never wait behind its instruction stream
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_q, kQTileU4 * 16);
      bulk_load(Qs, Qw + (bh * n_qb + qb) * kQTileU4, kQTileU4 * 16, bar_q);
      mbar_arrive_expect_tx(bar_k, kKTileU4 * 16);
      bulk_load(Ks, kt_src, kKTileU4 * 16, bar_k);
      mbar_arrive_expect_tx(bar_v, kKTileU4 * 16);
      bulk_load(Vs, vt_src, kKTileU4 * 16, bar_v);
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      issue_s();
      for (int kt = 0; kt < n_kt; ++kt) {
        const uint32_t ph = (uint32_t)(kt & 1);
        const bool more = kt + 1 < n_kt;
        mbar_wait(bar_s, ph);  // S(kt) done: the K buffer is free
        if (more) {
          mbar_arrive_expect_tx(bar_k, kKTileU4 * 16);
          bulk_load(Ks, kt_src + (int64_t)(kt + 1) * kKTileU4, kKTileU4 * 16, bar_k);
        }
        mbar_wait(bar_p, ph);  // P(kt) written, S(kt) and O(kt-1) read by every softmax thread
        mbar_wait(bar_v, ph);
        tc_fence_after();
        issue_pv();
        if (more) {  // S of the next tile queues right behind P V on the tensor pipe
          mbar_wait(bar_k, ph ^ 1u);
          issue_s();
        }
        mbar_wait(bar_o, ph);  // P V(kt) done: the V buffer is free
        if (more) {
          mbar_arrive_expect_tx(bar_v, kKTileU4 * 16);
          bulk_load(Vs, vt_src + (int64_t)(kt + 1) * kKTileU4, kKTileU4 * 16, bar_v);
        }
      }
    }
    __syncwarp();
  } else {
    // ===== softmax warps
    const uint32_t t_addr = ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32);
    for (int kt = 0; kt < n_kt; ++kt) {
      const int k0 = kt * kK + half * 32;
      const uint32_t ph = (uint32_t)(kt & 1);
      mbar_wait(bar_s, ph);
      tc_fence_after();
      float s[32];
      tmem_ld32(tmem_s + t_addr, s);
      if (k0 + 32 > T) {  // only the last tile has keys past the end
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + j >= T) s[j] = -INFINITY;
      }
      float tmax = s[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) tmax = fmaxf(tmax, s[j]);
      float* pm = pmax + ph * 2 * kQ;
      pm[half * kQ + row] = tmax;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      const float m_new = fmaxf(m_run, fmaxf(tmax, pm[(half ^ 1) * kQ + row]));
      const float corr = ex2_approx(m_run - m_new);  // 0 on the first tile (its first 32 keys are never masked)
      float psum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = ex2_approx(s[j] - m_new);
        psum += s[j];
      }
      l_run = fmaf(l_run, corr, psum);
      m_run = m_new;
      // P as bf16 hi | lo, K-major A operand: Ps[(split*8 + key/8)*128 + row]  (P V(kt-1) was waited for below)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(s + g * 8, hi, lo);
        Ps[(0 * 8 + half * 4 + g) * kQ + row] = hi;
        Ps[(1 * 8 + half * 4 + g) * kQ + row] = lo;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);
      mbar_wait(bar_o, ph);  // O(kt) is in TMEM
      tc_fence_after();
      float pv[32];
      tmem_ld32(tmem_o + t_addr, pv);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = fmaf(acc[j], corr, pv[j]);
    }
  }
  // combine the two halves' partial sums of the row
  __syncthreads();
  if (warp < kIssuerWarp) pmax[half * kQ + row] = l_run;
  __syncthreads();
  const float l_tot = warp < kIssuerWarp ? l_run + pmax[(half ^ 1) * kQ + row] : 1.f;
  const int tq = qb * kQ + row;
  if (warp < kIssuerWarp && tq < T) {
    const float inv = 1.0f / l_tot;
    if (out_split) {
      const int C = H * kD;
      const int64_t orow = (int64_t)b * T + tq;
      __nv_bfloat16* os = reinterpret_cast<__nv_bfloat16*>(o);
      uint4* oh = reinterpret_cast<uint4*>(os + orow * C + h * kD + half * 32);
      uint4* ol = reinterpret_cast<uint4*>(os + (o_bs + orow) * C + h * kD + half * 32);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = acc[g * 8 + e] * inv;
        uint4 hi, lo;
        split8(x, hi, lo);
        oh[g] = hi;
        ol[g] = lo;
      }
    } else {
      float* __restrict__ ob = o + (int64_t)b * o_bs + ((int64_t)h * kD + half * 32) * T;
#pragma unroll
      for (int j = 0; j < 32; ++j) ob[(int64_t)j * T + tq] = acc[j] * inv;
      if (lse && half == 0) lse[bh * T + tq] = (m_run + log2f(l_tot)) * 0.69314718055994531f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_s, 128);
}

// =====================================================================================================================
// Backward (no mask / RoPE / dropout, head dim 64), two tensor-core kernels over pre-split operand tiles:
//   s = scale q k^T, P = exp(s - lse), dP = dO V^T, delta_i = sum_d dO_id O_id, dS = P o (dP - delta)
//   dQ = scale dS K          dK = scale dS^T Q          dV = P^T dO
// dQ kernel : CTA = 128 queries, streams 128-key tiles: S = Q K^T and dP = dO V^T (M=128, N=128, K=64) into TMEM, the
//             threads form dS (2 threads per query row, lse / delta in registers) and re-stage it as bf16 hi|lo,
//             dQ += dS K (M=128, N=64, K=128) accumulates in TMEM over ALL tiles (no rescaling: lse is known).
// dKV kernel: CTA = 128 keys, streams 64-query tiles: S^T = K Q^T, dP^T = V dO^T (M=128 keys, N=64), the threads form
//             P^T and dS^T (lse / delta per COLUMN from shared memory), dV += P^T dO and dK += dS^T Q accumulate in TMEM.
// q arrives scaled by scale*log2(e) in the tiles: P = ex2(S - lse*log2e); dS is staged * scale (dQ kernel, K unscaled)
// or * ln2 (dKV kernel, against the scaled Q tiles).  bf16x3 everywhere.
constexpr int kBT = 128;                       // rows per owner / stream tile of the dQ kernel
constexpr int kT128U4 = 2 * 8 * 128, kT64U4 = 2 * 8 * 64;

// one (b, h, 64-row block): q (scaled), k, v, dO -> 128-row tile images (half of one) + q, dO 64-row tile images;
// delta = rowsum(dO o O), lse2 = lse * log2e (+inf for rows >= T: P = 0 there)
__global__ void __launch_bounds__(256)
attn64_bwd_prepare_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                          int64_t qkv_bs, const float* __restrict__ o, const float* __restrict__ d_o, int64_t o_bs,
                          const float* __restrict__ lse, uint4* __restrict__ Q128, uint4* __restrict__ K128,
                          uint4* __restrict__ V128, uint4* __restrict__ G128, uint4* __restrict__ Q64,
                          uint4* __restrict__ G64, float* __restrict__ delta, float* __restrict__ lse2, int T,
                          int n128, int n64, float qmul) {
  const int tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z, H = gridDim.y;
  const int t0 = tile * 64;
  const int64_t bh = (int64_t)b * H + h;
  const int64_t off_qkv = (int64_t)b * qkv_bs + (int64_t)h * kD * T, off_o = (int64_t)b * o_bs + (int64_t)h * kD * T;
  const int64_t t128 = (bh * n128 + (tile >> 1)) * kT128U4 + (tile & 1) * 64;
  const int64_t t64 = (bh * n64 + tile) * kT64U4;
  __shared__ float dsum[8][64];
  for (int item = threadIdx.x; item < 4 * 512; item += 256) {
    const int which = item >> 9, r = item & 511;
    const int d8 = r >> 6, row = r & 63, t = t0 + row;
    const float* __restrict__ src = which == 0 ? q + off_qkv : which == 1 ? k + off_qkv : which == 2 ? v + off_qkv
                                                                                                      : d_o + off_o;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = t < T ? src[(int64_t)(d8 * 8 + j) * T + t] : 0.f;
    if (which == 3) {  // delta partial over this thread's 8 features
      float acc = 0.f;
      if (t < T)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(x[j], o[off_o + (int64_t)(d8 * 8 + j) * T + t], acc);
      dsum[d8][row] = acc;
    }
    if (which == 0)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] *= qmul;
    uint4 hi, lo;
    split8(x, hi, lo);
    uint4* big = which == 0 ? Q128 : which == 1 ? K128 : which == 2 ? V128 : G128;
    big[t128 + (0 * 8 + d8) * 128 + row] = hi;
    big[t128 + (1 * 8 + d8) * 128 + row] = lo;
    if ((which == 0 || which == 3) && tile < n64) {
      uint4* sm = which == 0 ? Q64 : G64;
      sm[t64 + (0 * 8 + d8) * 64 + row] = hi;
      sm[t64 + (1 * 8 + d8) * 64 + row] = lo;
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int row = threadIdx.x, t = t0 + row;
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) a += dsum[j][row];
    const int64_t idx = bh * ((int64_t)n128 * 128) + t;  // padded to whole 128-row tiles
    if (t < n128 * 128) {
      delta[idx] = t < T ? a : 0.f;
      lse2[idx] = t < T ? lse[bh * T + t] * 1.4426950408889634f : INFINITY;
    }
  }
}

// generic bf16x3 MMA group: D (+)= A B over n_k K-steps; descriptors as (lo, hi) words, lo words advanced per K-step
__device__ __forceinline__ void mma_x3(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t a_kstep, uint32_t b_kstep,
                                       uint32_t a_lo_off, uint32_t b_lo_off, uint32_t idesc, int n_k, uint32_t acc0) {
  const uint32_t a_hi = (uint32_t)(ad >> 32), b_hi = (uint32_t)(bd >> 32);
  uint32_t ak = (uint32_t)ad, bk = (uint32_t)bd;
#pragma unroll 1
  for (int ks = 0; ks < n_k; ++ks, ak += a_kstep, bk += b_kstep) {
    umma_bf16_w(tmem_d, ak, a_hi, bk, b_hi, idesc, ks == 0 ? acc0 : 1u);
    umma_bf16_w(tmem_d, ak + a_lo_off, a_hi, bk, b_hi, idesc, 1u);  // lo * hi
    umma_bf16_w(tmem_d, ak, a_hi, bk + b_lo_off, b_hi, idesc, 1u);  // hi * lo
  }
}

__host__ __device__ constexpr uint32_t idesc_mn(int N, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
attn64_bwd_dq_kernel(const uint4* __restrict__ Q128, const uint4* __restrict__ K128, const uint4* __restrict__ V128,
                     const uint4* __restrict__ G128, const float* __restrict__ delta, const float* __restrict__ lse2,
                     float* __restrict__ dq, int64_t dq_bs, int T, int n128, float scale) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* Qs = reinterpret_cast<uint4*>(smem_raw);   // [2][8][128]
  uint4* Gs = Qs + kT128U4;                         // dO
  uint4* Ks = Gs + kT128U4;
  uint4* Vs = Ks + kT128U4;
  uint4* Ds = Vs + kT128U4;                         // dS [2][16 key groups][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ds + 2 * 16 * 128);
  uint64_t *bar_own = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_dp = bars + 4,
           *bar_ds = bars + 5, *bar_dq = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane;
  const int b = blockIdx.z, h = blockIdx.y, qb = blockIdx.x, H = gridDim.y;
  const int64_t bh = (int64_t)b * H + h;
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], i == 5 ? 256u : 1u);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_s = *tmem_slot, tm_dp = tm_s + 128, tm_dq = tm_s + 256;
  const uint4* kt_src = K128 + bh * n128 * kT128U4;
  const uint4* vt_src = V128 + bh * n128 * kT128U4;
  constexpr uint32_t TILE_B = kT128U4 * 16;

  if (warp == kIssuerWarp) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_own, 2 * TILE_B);
      bulk_load(Qs, Q128 + (bh * n128 + qb) * kT128U4, TILE_B, bar_own);
      bulk_load(Gs, G128 + (bh * n128 + qb) * kT128U4, TILE_B, bar_own);
      mbar_arrive_expect_tx(bar_k, TILE_B);
      bulk_load(Ks, kt_src, TILE_B, bar_k);
      mbar_arrive_expect_tx(bar_v, TILE_B);
      bulk_load(Vs, vt_src, TILE_B, bar_v);
      mbar_wait(bar_own, 0);
      const uint64_t q_d = make_desc(smem_u32(Qs), 128u, 8u), g_d = make_desc(smem_u32(Gs), 128u, 8u);
      const uint64_t k_d = make_desc(smem_u32(Ks), 128u, 8u), v_d = make_desc(smem_u32(Vs), 128u, 8u);
      const uint64_t ds_d = make_desc(smem_u32(Ds), 128u, 8u);           // A: K-major, 16 key groups
      const uint64_t kmn_d = make_desc(smem_u32(Ks), 8u, 128u);          // B: MN-major (N = d, K = keys)
      for (int kt = 0; kt < n128; ++kt) {
        const uint32_t ph = (uint32_t)(kt & 1);
        mbar_wait(bar_k, ph);
        tc_fence_after();
        mma_x3(tm_s, q_d, k_d, 256u, 256u, 8u * 128u, 8u * 128u, idesc_mn(128, false), 4, 0u);
        umma_commit(bar_s);
        mbar_wait(bar_v, ph);
        mma_x3(tm_dp, g_d, v_d, 256u, 256u, 8u * 128u, 8u * 128u, idesc_mn(128, false), 4, 0u);
        umma_commit(bar_dp);
        mbar_wait(bar_dp, ph);  // V is free
        if (kt + 1 < n128) {
          mbar_arrive_expect_tx(bar_v, TILE_B);
          bulk_load(Vs, vt_src + (int64_t)(kt + 1) * kT128U4, TILE_B, bar_v);
        }
        mbar_wait(bar_ds, ph);  // dS staged, S / dP read
        tc_fence_after();
        mma_x3(tm_dq, ds_d, kmn_d, 256u, 16u, 16u * 128u, 8u * 128u, idesc_mn(64, true), 8, kt == 0 ? 0u : 1u);
        umma_commit(bar_dq);
        mbar_wait(bar_dq, ph);  // K and dS buffers are free
        if (kt + 1 < n128) {
          mbar_arrive_expect_tx(bar_k, TILE_B);
          bulk_load(Ks, kt_src + (int64_t)(kt + 1) * kT128U4, TILE_B, bar_k);
        }
      }
    }
    __syncwarp();
  } else {
    const int64_t ridx = bh * ((int64_t)n128 * 128) + qb * 128 + row;
    const float my_lse = lse2[ridx], my_delta = delta[ridx];
    const uint32_t t_addr = ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64);
    for (int kt = 0; kt < n128; ++kt) {
      const uint32_t ph = (uint32_t)(kt & 1);
      const int k0 = kt * 128 + half * 64;
      mbar_wait(bar_s, ph);
      mbar_wait(bar_dp, ph);
      tc_fence_after();
      if (kt > 0) mbar_wait(bar_dq, ph ^ 1u);  // the previous tile's dQ MMAs have read the dS buffer
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float s[32], dp[32];
        tmem_ld32(tm_s + t_addr + 32u * c, s);
        tmem_ld32(tm_dp + t_addr + 32u * c, dp);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float p = (k0 + 32 * c + j < T) ? ex2_approx(s[j] - my_lse) : 0.f;
          s[j] = p * (dp[j] - my_delta) * scale;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 hi, lo;
          split8(s + g * 8, hi, lo);
          const int kg = half * 8 + c * 4 + g;
          Ds[(0 * 16 + kg) * 128 + row] = hi;
          Ds[(1 * 16 + kg) * 128 + row] = lo;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_ds);
    }
    mbar_wait(bar_dq, (uint32_t)((n128 - 1) & 1));
    tc_fence_after();
    const int tq = qb * 128 + row;
    float acc[32];
    tmem_ld32(tm_dq + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32), acc);
    if (tq < T) {
      float* __restrict__ ob = dq + (int64_t)b * dq_bs + ((int64_t)h * kD + half * 32) * T;
#pragma unroll
      for (int j = 0; j < 32; ++j) ob[(int64_t)j * T + tq] = acc[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm_s, 512);
}

__global__ void __launch_bounds__(kThreads, 1)
attn64_bwd_dkv_kernel(const uint4* __restrict__ K128, const uint4* __restrict__ V128, const uint4* __restrict__ Q64,
                      const uint4* __restrict__ G64, const float* __restrict__ delta, const float* __restrict__ lse2,
                      float* __restrict__ dk, float* __restrict__ dv, int64_t d_bs, int T, int n128, int n64) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* Ks = reinterpret_cast<uint4*>(smem_raw);   // [2][8][128] owner keys
  uint4* Vs = Ks + kT128U4;
  uint4* Qs = Vs + kT128U4;                         // [2][8][64] streamed queries
  uint4* Gs = Qs + kT64U4;                          // dO
  uint4* Ps = Gs + kT64U4;                          // P^T  [2][8 query groups][128 keys]
  uint4* Ds = Ps + 2 * 8 * 128;                     // dS^T
  float* col = reinterpret_cast<float*>(Ds + 2 * 8 * 128);  // [2 parity][lse2 | delta][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(col + 2 * 2 * 64);
  uint64_t *bar_own = bars, *bar_q = bars + 1, *bar_s = bars + 2, *bar_dp = bars + 3, *bar_pd = bars + 4,
           *bar_acc = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2, row = quad * 32 + lane;
  const int b = blockIdx.z, h = blockIdx.y, kb = blockIdx.x, H = gridDim.y;
  const int64_t bh = (int64_t)b * H + h;
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(&bars[i], i == 4 ? 256u : 1u);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm_s = *tmem_slot, tm_dp = tm_s + 64, tm_dv = tm_s + 128, tm_dk = tm_s + 192;
  const uint4* q_src = Q64 + bh * n64 * kT64U4;
  const uint4* g_src = G64 + bh * n64 * kT64U4;
  const float* lse_src = lse2 + bh * ((int64_t)n128 * 128);
  const float* del_src = delta + bh * ((int64_t)n128 * 128);
  constexpr uint32_t OWN_B = kT128U4 * 16, STR_B = kT64U4 * 16;

  if (warp == kIssuerWarp) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_own, 2 * OWN_B);
      bulk_load(Ks, K128 + (bh * n128 + kb) * kT128U4, OWN_B, bar_own);
      bulk_load(Vs, V128 + (bh * n128 + kb) * kT128U4, OWN_B, bar_own);
      mbar_arrive_expect_tx(bar_q, 2 * STR_B);
      bulk_load(Qs, q_src, STR_B, bar_q);
      bulk_load(Gs, g_src, STR_B, bar_q);
      mbar_wait(bar_own, 0);
      const uint64_t k_d = make_desc(smem_u32(Ks), 128u, 8u), v_d = make_desc(smem_u32(Vs), 128u, 8u);
      const uint64_t q_d = make_desc(smem_u32(Qs), 64u, 8u), g_d = make_desc(smem_u32(Gs), 64u, 8u);
      const uint64_t p_d = make_desc(smem_u32(Ps), 128u, 8u), ds_d = make_desc(smem_u32(Ds), 128u, 8u);
      const uint64_t qmn_d = make_desc(smem_u32(Qs), 8u, 64u), gmn_d = make_desc(smem_u32(Gs), 8u, 64u);
      for (int qt = 0; qt < n64; ++qt) {
        const uint32_t ph = (uint32_t)(qt & 1);
        mbar_wait(bar_q, ph);
        tc_fence_after();
        mma_x3(tm_s, k_d, q_d, 256u, 128u, 8u * 128u, 8u * 64u, idesc_mn(64, false), 4, 0u);
        umma_commit(bar_s);
        mma_x3(tm_dp, v_d, g_d, 256u, 128u, 8u * 128u, 8u * 64u, idesc_mn(64, false), 4, 0u);
        umma_commit(bar_dp);
        mbar_wait(bar_pd, ph);  // P^T and dS^T staged, S^T / dP^T read
        tc_fence_after();
        mma_x3(tm_dv, p_d, gmn_d, 256u, 16u, 8u * 128u, 8u * 64u, idesc_mn(64, true), 4, qt == 0 ? 0u : 1u);
        mma_x3(tm_dk, ds_d, qmn_d, 256u, 16u, 8u * 128u, 8u * 64u, idesc_mn(64, true), 4, qt == 0 ? 0u : 1u);
        umma_commit(bar_acc);
        mbar_wait(bar_acc, ph);  // Q / dO tiles and the P^T / dS^T buffers are free
        if (qt + 1 < n64) {
          mbar_arrive_expect_tx(bar_q, 2 * STR_B);
          bulk_load(Qs, q_src + (int64_t)(qt + 1) * kT64U4, STR_B, bar_q);
          bulk_load(Gs, g_src + (int64_t)(qt + 1) * kT64U4, STR_B, bar_q);
        }
      }
    }
    __syncwarp();
  } else {
    const uint32_t t_addr = ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 32);
    for (int qt = 0; qt < n64; ++qt) {
      const uint32_t ph = (uint32_t)(qt & 1);
      float* cs = col + ph * 128;
      if (tid < 64) {  // per-column (query) constants of this tile
        cs[tid] = lse_src[qt * 64 + tid];
        cs[64 + tid] = del_src[qt * 64 + tid];
      }
      mbar_wait(bar_s, ph);
      mbar_wait(bar_dp, ph);
      tc_fence_after();
      if (qt > 0) mbar_wait(bar_acc, ph ^ 1u);  // previous dV / dK MMAs have read the staging buffers
      asm volatile("bar.sync 1, 256;" ::: "memory");  // column constants visible (the 8 softmax warps only)
      float s[32], dp[32];
      tmem_ld32(tm_s + t_addr, s);
      tmem_ld32(tm_dp + t_addr, dp);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float p = ex2_approx(s[j] - cs[half * 32 + j]);  // lse2 = +inf for padded queries: p = 0
        dp[j] = p * (dp[j] - cs[64 + half * 32 + j]) * 0.69314718055994531f;
        s[j] = p;
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(s + g * 8, hi, lo);
        Ps[(0 * 8 + half * 4 + g) * 128 + row] = hi;
        Ps[(1 * 8 + half * 4 + g) * 128 + row] = lo;
        split8(dp + g * 8, hi, lo);
        Ds[(0 * 8 + half * 4 + g) * 128 + row] = hi;
        Ds[(1 * 8 + half * 4 + g) * 128 + row] = lo;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pd);
    }
    mbar_wait(bar_acc, (uint32_t)((n64 - 1) & 1));
    tc_fence_after();
    const int tk = kb * 128 + row;
    float a[32];
    tmem_ld32(tm_dv + t_addr, a);
    if (tk < T) {
      float* __restrict__ ob = dv + (int64_t)b * d_bs + ((int64_t)h * kD + half * 32) * T;
#pragma unroll
      for (int j = 0; j < 32; ++j) ob[(int64_t)j * T + tk] = a[j];
    }
    tmem_ld32(tm_dk + t_addr, a);
    if (tk < T) {
      float* __restrict__ ob = dk + (int64_t)b * d_bs + ((int64_t)h * kD + half * 32) * T;
#pragma unroll
      for (int j = 0; j < 32; ++j) ob[(int64_t)j * T + tk] = a[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm_s, 256);
}

struct Plan {
  int n_qb, n_kt;
  int64_t q_u4, kv_u4;
};
Plan make_plan(int B, int H, int T) {
  Plan p;
  p.n_qb = cdiv(T, kQ);
  p.n_kt = cdiv(T, kK);
  p.q_u4 = (int64_t)B * H * p.n_qb * kQTileU4;
  p.kv_u4 = (int64_t)B * H * p.n_kt * kKTileU4;
  return p;
}

template <bool TM, bool OUT_SPLIT>
int launch(const float* q, const float* k, const float* v, int64_t bs, float* o, int64_t o_bs, int B, int H, int T,
           float scale, float* lse, void* workspace, cudaStream_t st) {
  const Plan p = make_plan(B, H, T);
  uint4* Qw = reinterpret_cast<uint4*>(workspace);
  uint4* Kw = Qw + p.q_u4;
  uint4* Vw = Kw + p.kv_u4;
  attn64_prepare_kernel<TM><<<dim3(2 * p.n_qb, H, B), 256, 0, st>>>(q, k, v, bs, Qw, Kw, Vw, T, p.n_qb, p.n_kt,
                                                                    scale * 1.4426950408889634f);
  const size_t smem = (size_t)(2 * kQTileU4 + 2 * kKTileU4) * 16 + 2 * 2 * kQ * sizeof(float) + 128;
  cudaFuncSetAttribute(attn64_kernel_impl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn64_kernel_impl<<<dim3(p.n_qb, H, B), kThreads, smem, st>>>(Qw, Kw, Vw, o, o_bs, T, p.n_qb, p.n_kt, lse,
                                                                 OUT_SPLIT);
  return 0;
}

}  // namespace
}  // namespace sty

using namespace sty;

extern "C" int64_t sty_attention64_workspace_bytes(int B, int H, int T) {
  if (B <= 0 || H <= 0 || T <= 0) return 0;
  const Plan p = make_plan(B, H, T);
  return (p.q_u4 + 2 * p.kv_u4) * 16;
}

extern "C" int sty_attention64_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs, float* o,
                                   int64_t o_bs, int B, int H, int T, float scale, float* lse, void* workspace,
                                   sty_stream_t stream) {
  STY_REQUIRE(q && k && v && o && workspace && B > 0 && H > 0 && T > 0 && H <= 65535 && B <= 65535,
              "attention64: bad argument");
  STY_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "attention64: workspace must be 16-byte aligned");
  launch<false, false>(q, k, v, qkv_bs, o, o_bs, B, H, T, scale, lse, workspace, as_stream(stream));
  STY_CHECK_LAUNCH("attention64");
  return STY_OK;
}

extern "C" int sty_attention64_tokens_fwd(const float* qkv, int64_t ld, void* out_split, int64_t M_pad, int B, int H,
                                          int T, float scale, void* workspace, sty_stream_t stream) {
  STY_REQUIRE(qkv && out_split && workspace && B > 0 && H > 0 && T > 0 && ld >= 3 * H * kD && (ld & 3) == 0 &&
                  M_pad >= (int64_t)B * T && H <= 65535 && B <= 65535,
              "attention64_tokens: bad argument");
  STY_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0,
              "attention64_tokens: workspace / qkv must be 16-byte aligned");
  launch<true, true>(qkv, qkv + H * kD, qkv + 2 * H * kD, ld, reinterpret_cast<float*>(out_split), M_pad, B, H, T,
                     scale, nullptr, workspace, as_stream(stream));
  STY_CHECK_LAUNCH("attention64_tokens");
  return STY_OK;
}

extern "C" int64_t sty_attention64_bwd_workspace_bytes(int B, int H, int T) {
  if (B <= 0 || H <= 0 || T <= 0) return 0;
  const int64_t n128 = cdiv(T, 128), n64 = cdiv(T, 64), bh = (int64_t)B * H;
  return (bh * n128 * kT128U4 * 4 + bh * n64 * kT64U4 * 2) * 16 + bh * n128 * 128 * 2 * (int64_t)sizeof(float);
}

extern "C" int sty_attention64_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                                   const float* d_o, int64_t o_bs, const float* lse, float* dq, float* dk, float* dv,
                                   int64_t dqkv_bs, int B, int H, int T, float scale, void* workspace,
                                   sty_stream_t stream) {
  STY_REQUIRE(q && k && v && o && d_o && lse && dq && dk && dv && workspace && B > 0 && H > 0 && T >= 64 &&
                  H <= 65535 && B <= 65535,
              "attention64_bwd: bad argument");
  STY_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "attention64_bwd: workspace must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int n128 = cdiv(T, 128), n64 = cdiv(T, 64);
  const int64_t bh = (int64_t)B * H;
  uint4* Q128 = reinterpret_cast<uint4*>(workspace);
  uint4* K128 = Q128 + bh * n128 * kT128U4;
  uint4* V128 = K128 + bh * n128 * kT128U4;
  uint4* G128 = V128 + bh * n128 * kT128U4;
  uint4* Q64 = G128 + bh * n128 * kT128U4;
  uint4* G64 = Q64 + bh * n64 * kT64U4;
  float* delta = reinterpret_cast<float*>(G64 + bh * n64 * kT64U4);
  float* lse2 = delta + bh * n128 * 128;
  attn64_bwd_prepare_kernel<<<dim3(2 * n128, H, B), 256, 0, st>>>(q, k, v, qkv_bs, o, d_o, o_bs, lse, Q128, K128, V128,
                                                                  G128, Q64, G64, delta, lse2, T, n128, n64,
                                                                  scale * 1.4426950408889634f);
  {
    const size_t smem = (size_t)(4 * kT128U4 + 2 * 16 * 128) * 16 + 128;
    cudaFuncSetAttribute(attn64_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn64_bwd_dq_kernel<<<dim3(n128, H, B), kThreads, smem, st>>>(Q128, K128, V128, G128, delta, lse2, dq, dqkv_bs, T,
                                                                   n128, scale);
  }
  {
    const size_t smem = (size_t)(2 * kT128U4 + 2 * kT64U4 + 2 * 2 * 8 * 128) * 16 + 2 * 2 * 64 * sizeof(float) + 128;
    cudaFuncSetAttribute(attn64_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attn64_bwd_dkv_kernel<<<dim3(n128, H, B), kThreads, smem, st>>>(K128, V128, Q64, G64, delta, lse2, dk, dv, dqkv_bs,
                                                                    T, n128, n64);
  }
  STY_CHECK_LAUNCH("attention64_bwd");
  return STY_OK;
}
