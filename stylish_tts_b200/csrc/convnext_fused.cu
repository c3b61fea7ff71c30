// GeneratorConvNeXtBlock (conv_next.py:80-93, GRN :7-18) at the vocoder's output rate (C = 32 channels,
// 4C = 128 hidden, T ~ 60 000 steps) WITHOUT materialising the 4C-wide intermediate:
//
//   pass 1  x --TMA--> [dwconv7 + LN_C + AdaLN] --tcgen05--> h = Snake(W1 x^ + b1)      sum_t h^2 only
//   (GRN)   gs[b,j] = 1 + gamma_j * ||h_j|| / (mean_j ||h_j|| + 1e-6)                     (sty_grn_scale_fwd)
//   pass 2  x --TMA--> [same front] --tcgen05--> h --(* gs, bf16 hi|lo, smem)--tcgen05--> W2 h + b2' + x --> y
//
// HBM traffic per block: read x twice, write y once = 3 u (u = one 32-channel tensor) instead of 11 u for the
// two-kernel version that stores h (SURVEY 8d budgets 3 u).  What bounds the kernel is instruction issue in the
// Snake epilogue (128 x T values, twice), so the layout is chosen to minimise instructions per value:
//
//   * both GEMMs are computed TRANSPOSED: D1^T[j, t] = W1^T x^^T and D2^T[co, t] = W2^T h^T.  TMEM lanes are
//     then CHANNELS and TMEM columns are time steps, so an epilogue thread owns one hidden channel j: bias,
//     Snake alpha, GRN scale are registers (no shared-memory parameter loads), sum_t h^2 is a private register
//     (no shuffle butterflies, one atomic per thread and batch row), and 8 consecutive time steps of one channel
//     are exactly one 16-byte row of the MN-major B operand of the second GEMM (conflict-free 512-byte warp stores).
//   * Snake as  h = w - (1/2a) cos(2a w - 1),  w = v + 1/2a   (= v + sin^2(a v)/a): FADD, FFMA, [FMUL], MUFU.COS,
//     FFMA per value.
//   * the raw fp32 tile arrives by ONE cp.async.bulk.tensor per 128 steps (halo included, zero fill at the ends),
//     3 stages ahead; the front threads own two adjacent time steps (sliding window: 5 shared loads per channel
//     for 14 FMAs).
//
// Warp roles: 0 TMA loader | 1 MMA issuer (elect.sync) | 4-7 front (two groups of 64 threads alternate tiles) |
// 8-15 Snake epilogue (lane quadrant = warp % 4, column half = (warp-8)/4) | 16-17 output epilogue (pass 2).
// bf16x3 split precision as everywhere (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM).
#include <string.h>

#include "tma.cuh"

extern "C" int sty_grn_scale_fwd(const float* sumsq, const float* gamma, float* scale, int B, int J,
                                 sty_stream_t stream);

namespace sty {
namespace {

constexpr int kC = 32, kJ = 128, kMT = 128;
constexpr int kRawW = 136;  // box: steps t0-4 .. t0+131 (the box must start on a 16-byte boundary)
constexpr int kRawStages = 4;  // pass 2 keeps a stage until the output epilogue has read the residual from it
constexpr int kRawFloats = kC * kRawW;
constexpr int kMmaWarp = 1, kFrontWarp0 = 4, kEpi1Warp0 = 8, kEpi2Warp0 = 16;
constexpr int kThreads1 = 16 * 32, kThreads2 = 18 * 32;

// shared-memory carve-up (bytes)
constexpr int kOffRaw = 0;
constexpr int kOffW1 = kOffRaw + kRawStages * kRawFloats * 4;  // [2 split][4 c8][128 j] x 16 B
constexpr int kOffX = kOffW1 + 2 * 4 * kJ * 16;                // 2 stages x [2][4][128 t] x 16 B
constexpr int kOffPrm = kOffX + 2 * 2 * 4 * kMT * 16;          // dw taps+bias [32][8], gamma|beta per group [2][64]
constexpr int kOffBars = kOffPrm + (kC * 8 + 2 * 64) * 4;
constexpr int kOffW2 = kOffBars + 256;                         // pass 2: [2][16 j8][64 co] x 16 B
constexpr int kOffH = kOffW2 + 2 * 16 * 64 * 16;               // pass 2: [2][16 tg][128 j] x 16 B
constexpr int kSmem1 = kOffW2, kSmem2 = kOffH + 2 * 16 * kJ * 16;

struct Args {
  const float* x;
  int64_t x_bs, x_cs;
  float* y;
  int64_t y_bs, y_cs;
  const float *dw_w, *dw_b, *gb;  // (32,7), (32), gamma|beta rows
  int64_t gb_bs;
  float eps;
  const uint4 *w1s, *w2s;  // bf16 hi|lo packs: [2][4][128] and [2][16][32] 16-byte units
  const float *b1, *alpha, *b2;
  float* sumsq;      // (B,128)
  const float* gs;   // (B,128)
  int B, T, tiles_per_b;
};

template <int PASS>
__global__ void __launch_bounds__(PASS == 1 ? kThreads1 : kThreads2, 1)
convnext_fused_kernel(const Args p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* raw0 = reinterpret_cast<float*>(smem + kOffRaw);
  uint4* W1s = reinterpret_cast<uint4*>(smem + kOffW1);
  uint4* Xs = reinterpret_cast<uint4*>(smem + kOffX);
  float* dwp = reinterpret_cast<float*>(smem + kOffPrm);  // [32][8]
  float* gbs = dwp + kC * 8;                              // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* raw_full = bars;        // [4]
  uint64_t* raw_empty = bars + 4;   // [4]
  uint64_t* x_full = bars + 8;      // [2]
  uint64_t* x_empty = bars + 10;    // [2]
  uint64_t* acc1_full = bars + 12;  // [2]
  uint64_t* acc1_empty = bars + 14; // [2]
  uint64_t* h_full = bars + 16;     // [2] half tiles of 64 steps: GEMM 2 of one half overlaps the Snake of the next
  uint64_t* h_empty = bars + 18;    // [2]
  uint64_t* acc2_full = bars + 20;  // [2]
  uint64_t* acc2_ready = bars + 22; // [2] accumulator stage initialised with the residual (see the output epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint4* W2s = reinterpret_cast<uint4*>(smem + kOffW2);
  uint4* Hs = reinterpret_cast<uint4*>(smem + kOffH);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NT = PASS == 1 ? kThreads1 : kThreads2;

  if (warp == kMmaWarp) tmem_alloc(tmem_slot, PASS == 1 ? 256u : 512u);
  if (tid == 0) {
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(&raw_full[i], 1);
      mbar_init(&raw_empty[i], PASS == 1 ? 64 : 128);  // front group (+ output epilogue: residual)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 64);
      mbar_init(&x_empty[i], 1);
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], 256);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_ready[i], 64);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 256);
      mbar_init(&h_empty[i], 1);
    }
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * 4 * kJ; i += NT) W1s[i] = p.w1s[i];
  for (int i = tid; i < kC * 8; i += NT) {
    const int c = i >> 3, k = i & 7;
    dwp[i] = k < 7 ? p.dw_w[c * 7 + k] : p.dw_b[c];
  }
  if (PASS == 2) {
    for (int i = tid; i < 2 * 16 * 64; i += NT) {
      const int co = i & 63, blk = i >> 6;  // blk = split*16 + j8
      W2s[i] = co < 32 ? p.w2s[blk * 32 + co] : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = (int64_t)p.B * p.tiles_per_b;
  const int tile_begin = (int)((n_tiles * blockIdx.x) / gridDim.x);
  const int tile_end = (int)((n_tiles * (blockIdx.x + 1)) / gridDim.x);
  const int n_local = tile_end - tile_begin;

  if (warp == 0) {
    // =========================== TMA loader
    if (elect_one()) {
      prefetch_tensormap(&tmap);
      for (int it = 0; it < n_local; ++it) {
        const int tile = tile_begin + it;
        const int b = tile / p.tiles_per_b, t0 = (tile - b * p.tiles_per_b) * kMT;
        const uint32_t r = (uint32_t)it % kRawStages, ph = ((uint32_t)it / kRawStages) & 1u;
        mbar_wait_sleep(&raw_empty[r], ph ^ 1u);
        mbar_arrive_expect_tx(&raw_full[r], kRawFloats * 4);
        tma_load_3d(raw0 + r * kRawFloats, &tmap, &raw_full[r], t0 - 4, 0, b);
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer: GEMM 1 of tile it+1 is issued before GEMM 2 of tile it
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kMT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                            ((uint32_t)(64 >> 4) << 24);  // M = 64 (32 output channels used), N = 64 steps, B MN-major
    const uint64_t a1d = make_desc(smem_u32(W1s), 128u, 8u);
    const uint64_t a2d = make_desc(smem_u32(W2s), 64u, 8u);
    const uint64_t b2d = make_desc(smem_u32(Hs), 8u, 128u);
    auto gemm1 = [&](int it) {
      const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
      mbar_wait(&x_full[s], ph);
      mbar_wait(&acc1_empty[s], ph ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t b1d = make_desc(smem_u32(Xs + s * (2 * 4 * kMT)), 128u, 8u);
        const uint32_t ah = (uint32_t)(a1d >> 32), bh = (uint32_t)(b1d >> 32);
        const uint32_t al = (uint32_t)a1d, bl = (uint32_t)b1d;
        const uint32_t d = tmem_base + s * 128u;
#pragma unroll
        for (uint32_t ks = 0; ks < 2; ++ks) {
          const uint32_t ak = al + ks * 256u, bk = bl + ks * 256u;
          umma_bf16_w(d, ak, ah, bk, bh, idesc1, ks);            // W hi * x hi
          umma_bf16_w(d, ak + 512u, ah, bk, bh, idesc1, 1u);     // W lo * x hi
          umma_bf16_w(d, ak, ah, bk + 512u, bh, idesc1, 1u);     // W hi * x lo
        }
        umma_commit(&x_empty[s]);
        umma_commit(&acc1_full[s]);
      }
      __syncwarp();
    };
    auto gemm2 = [&](int it) {
      const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
      const uint32_t ah = (uint32_t)(a2d >> 32), bh = (uint32_t)(b2d >> 32);
      const uint32_t al = (uint32_t)a2d;
#pragma unroll 1
      for (uint32_t hh = 0; hh < 2; ++hh) {  // the two 64-step halves of the tile
        mbar_wait(&h_full[hh], (uint32_t)it & 1u);
        if (hh == 0) mbar_wait(&acc2_ready[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bl = (uint32_t)b2d + hh * 2048u;  // half stage: [2 split][8 tg][128 j] 16-byte units
          const uint32_t d = tmem_base + 256u + s * 128u + hh * 64u;
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks) {
            const uint32_t ak = al + ks * 128u, bk = bl + ks * 16u;
            umma_bf16_w(d, ak, ah, bk, bh, idesc2, 1u);            // W2 hi * h hi (on top of the residual)
            umma_bf16_w(d, ak + 1024u, ah, bk, bh, idesc2, 1u);    // W2 lo * h hi
            umma_bf16_w(d, ak, ah, bk + 1024u, bh, idesc2, 1u);    // W2 hi * h lo
          }
          umma_commit(&h_empty[hh]);
          if (hh == 1) umma_commit(&acc2_full[s]);
        }
        __syncwarp();
      }
    };
    if (n_local > 0) gemm1(0);
    for (int it = 0; it < n_local; ++it) {
      if (it + 1 < n_local) gemm1(it + 1);
      if (PASS == 2) gemm2(it);
    }
  } else if (warp >= kFrontWarp0 && warp < kFrontWarp0 + 4) {
    // =========================== front: depthwise k7 + LayerNorm over channels + adaptive affine
    const int g = (warp - kFrontWarp0) >> 1;            // group: tiles it = g, g+2, ...; operand stage g
    const int u = ((warp - kFrontWarp0) & 1) * 32 + lane;  // rows 2u, 2u+1 of the tile
    float* gbg = gbs + g * 64;
    int cur_b = -1;
    for (int it = g; it < n_local; it += 2) {
      const int tile = tile_begin + it;
      const int b = tile / p.tiles_per_b, t0 = (tile - b * p.tiles_per_b) * kMT;
      if (b != cur_b) {  // gamma | beta of this batch row (group-local barrier: ids 1, 2)
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
        gbg[u] = p.gb[(int64_t)b * p.gb_bs + u];
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory");
        cur_b = b;
      }
      const uint32_t r = (uint32_t)it % kRawStages, rph = ((uint32_t)it / kRawStages) & 1u;
      mbar_wait(&raw_full[r], rph);
      const float* raw = raw0 + r * kRawFloats + 2 * u;  // row 2u reads columns 2u+1 .. 2u+7
      float dA[kC], dB[kC];
      float sA = 0.f, sB = 0.f;
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const float* xr = raw + c * kRawW;
        const float v0 = xr[1];
        const float2 p1 = *reinterpret_cast<const float2*>(xr + 2);
        const float2 p2 = *reinterpret_cast<const float2*>(xr + 4);
        const float2 p3 = *reinterpret_cast<const float2*>(xr + 6);
        const float2 p4 = *reinterpret_cast<const float2*>(xr + 8);
        const float4 wa = *reinterpret_cast<const float4*>(dwp + c * 8);
        const float4 wb = *reinterpret_cast<const float4*>(dwp + c * 8 + 4);  // w4 w5 w6 bias
        float a = wb.w, bb = wb.w;
        a = fmaf(wa.x, v0, a);     bb = fmaf(wa.x, p1.x, bb);
        a = fmaf(wa.y, p1.x, a);   bb = fmaf(wa.y, p1.y, bb);
        a = fmaf(wa.z, p1.y, a);   bb = fmaf(wa.z, p2.x, bb);
        a = fmaf(wa.w, p2.x, a);   bb = fmaf(wa.w, p2.y, bb);
        a = fmaf(wb.x, p2.y, a);   bb = fmaf(wb.x, p3.x, bb);
        a = fmaf(wb.y, p3.x, a);   bb = fmaf(wb.y, p3.y, bb);
        a = fmaf(wb.z, p3.y, a);   bb = fmaf(wb.z, p4.x, bb);
        dA[c] = a;
        dB[c] = bb;
        sA += a;
        sB += bb;
      }
      mbar_arrive(&raw_empty[r]);  // raw tile read
      const float mA = sA * (1.0f / kC), mB = sB * (1.0f / kC);
      float qA = 0.f, qB = 0.f;
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const float ea = dA[c] - mA, eb = dB[c] - mB;
        qA = fmaf(ea, ea, qA);
        qB = fmaf(eb, eb, qB);
      }
      const bool okA = t0 + 2 * u < p.T, okB = t0 + 2 * u + 1 < p.T;
      // rows past the end are zeros (like the two-kernel path), folded into the normalisation factor
      const float rA = okA ? rsqrtf(qA * (1.0f / kC) + p.eps) : 0.f;
      const float rB = okB ? rsqrtf(qB * (1.0f / kC) + p.eps) : 0.f;
      mbar_wait(&x_empty[g], (((uint32_t)it >> 1) & 1u) ^ 1u);
      uint4* Xg = Xs + g * (2 * 4 * kMT) + 2 * u;
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t hA[4], lA[4], hB[4], lB[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c8 * 8 + 2 * j;
          const float2 gm = *reinterpret_cast<const float2*>(gbg + c);
          const float2 bt = *reinterpret_cast<const float2*>(gbg + 32 + c);
          const float s0 = (1.0f + gm.x), s1 = (1.0f + gm.y);
          const float a0 = okA ? fmaf(s0 * rA, dA[c] - mA, bt.x) : 0.f;
          const float a1 = okA ? fmaf(s1 * rA, dA[c + 1] - mA, bt.y) : 0.f;
          const float b0 = okB ? fmaf(s0 * rB, dB[c] - mB, bt.x) : 0.f;
          const float b1 = okB ? fmaf(s1 * rB, dB[c + 1] - mB, bt.y) : 0.f;
          hA[j] = pack_bf16(a0, a1);
          lA[j] = pack_bf16(a0 - __uint_as_float(hA[j] << 16), a1 - __uint_as_float(hA[j] & 0xffff0000u));
          hB[j] = pack_bf16(b0, b1);
          lB[j] = pack_bf16(b0 - __uint_as_float(hB[j] << 16), b1 - __uint_as_float(hB[j] & 0xffff0000u));
        }
        Xg[(0 * 4 + c8) * kMT] = make_uint4(hA[0], hA[1], hA[2], hA[3]);
        Xg[(0 * 4 + c8) * kMT + 1] = make_uint4(hB[0], hB[1], hB[2], hB[3]);
        Xg[(1 * 4 + c8) * kMT] = make_uint4(lA[0], lA[1], lA[2], lA[3]);
        Xg[(1 * 4 + c8) * kMT + 1] = make_uint4(lB[0], lB[1], lB[2], lB[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&x_full[g]);
    }
  } else if (warp >= kEpi1Warp0 && warp < kEpi1Warp0 + 8) {
    // =========================== Snake epilogue: thread = hidden channel j, columns = time steps
    const int q = warp & 3, half = (warp - kEpi1Warp0) >> 2;
    const int j = q * 32 + lane;
    const float al = p.alpha[j], ia = 1.0f / al;
    const float bp = p.b1[j] + 0.5f * ia, a2 = 2.0f * al, nh = -0.5f * ia;
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32);
    float ss = 0.f, gsv = 1.f;
    int cur_b = -1;
    for (int it = 0; it < n_local; ++it) {
      const int tile = tile_begin + it;
      const int b = tile / p.tiles_per_b, t0 = (tile - b * p.tiles_per_b) * kMT;
      if (b != cur_b) {
        if (PASS == 1) {
          if (cur_b >= 0) atomicAdd(p.sumsq + (int64_t)cur_b * kJ + j, ss);
          ss = 0.f;
        } else {
          gsv = p.gs[(int64_t)b * kJ + j];
        }
        cur_b = b;
      }
      const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
      mbar_wait_sleep(&acc1_full[s], ph);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // this warp's 32 columns of each 64-step half of the tile
        const int col0 = hh * 64 + half * 32;
        const int valid = p.T - t0 - col0;  // columns [0, valid) are real time steps
        float r[32];
        tmem_ld32(tmem_base + s * 128u + lane_addr + (uint32_t)(hh * 64), r);
        if (hh == 1) {  // both chunks are in registers: the accumulator stage may be overwritten
          tc_fence_before();
          mbar_arrive(&acc1_empty[s]);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float w = r[i] + bp;
          const float c = __cosf(fmaf(w, a2, -1.0f));
          r[i] = fmaf(c, nh, w);
        }
        if (PASS == 1) {
          if (valid >= 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) ss = fmaf(r[i], r[i], ss);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < valid) ss = fmaf(r[i], r[i], ss);
          }
        } else {
          uint4* Hh = Hs + hh * 2048;  // half stage [2 split][8 tg][128 j]
          mbar_wait(&h_empty[hh], ((uint32_t)it & 1u) ^ 1u);  // GEMM 2 of the previous tile has read this half
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            uint32_t hv[4], lv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v0 = r[g8 * 8 + 2 * e] * gsv, v1 = r[g8 * 8 + 2 * e + 1] * gsv;
              hv[e] = pack_bf16(v0, v1);
              lv[e] = pack_bf16(v0 - __uint_as_float(hv[e] << 16), v1 - __uint_as_float(hv[e] & 0xffff0000u));
            }
            const int tg = half * 4 + g8;
            Hh[(0 * 8 + tg) * kJ + j] = make_uint4(hv[0], hv[1], hv[2], hv[3]);
            Hh[(1 * 8 + tg) * kJ + j] = make_uint4(lv[0], lv[1], lv[2], lv[3]);
          }
          fence_proxy_async_smem();
          mbar_arrive(&h_full[hh]);
        }
      }
    }
    if (PASS == 1 && cur_b >= 0) atomicAdd(p.sumsq + (int64_t)cur_b * kJ + j, ss);
  } else if (PASS == 2 && warp >= kEpi2Warp0) {
    // =========================== output epilogue: D2^T (M = 64): rows 0-15 -> lanes 0-15, rows 16-31 -> lanes 32-47
    const int q = warp & 3;  // 16 -> 0, 17 -> 1
    const int co = q * 16 + (lane & 15);
    const bool has = lane < 16;
    const float b2 = p.b2[co];
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    // The residual is ADDED BY THE TENSOR CORE: this role initialises accumulator stage it & 1 with the block input
    // of tile it (from the raw TMA stage, as soon as it has landed), GEMM 2 then accumulates on top of it.  The raw
    // stage is released two tiles before the output of its tile is written, so the loader stays ahead.
    auto prefill = [&](int it2) {
      if (it2 >= n_local) return;
      const uint32_t s2 = (uint32_t)it2 & 1u, rs = (uint32_t)it2 % kRawStages, rph = ((uint32_t)it2 / kRawStages) & 1u;
      mbar_wait(&raw_full[rs], rph);
      const float* xr = raw0 + rs * kRawFloats + co * kRawW + 4;  // column 4 = step t0 (zeros past the end)
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        float r[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(xr + ch * 32 + 4 * i);
          r[4 * i] = v.x, r[4 * i + 1] = v.y, r[4 * i + 2] = v.z, r[4 * i + 3] = v.w;
        }
        tmem_st32(tmem_base + 256u + s2 * 128u + lane_addr + (uint32_t)(ch * 32), r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&acc2_ready[s2]);
      mbar_arrive(&raw_empty[rs]);  // second reader of the raw stage (after the front group)
    };
    prefill(0);
    prefill(1);
    for (int it = 0; it < n_local; ++it) {
      const int tile = tile_begin + it;
      const int b = tile / p.tiles_per_b, t0 = (tile - b * p.tiles_per_b) * kMT;
      const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
      float* __restrict__ yr = p.y + (int64_t)b * p.y_bs + (int64_t)co * p.y_cs + t0;
      mbar_wait_sleep(&acc2_full[s], ph);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int tc = t0 + ch * 32;
        float r[32];
        tmem_ld32(tmem_base + 256u + s * 128u + lane_addr + (uint32_t)(ch * 32), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (has && (tc + 4 * i < p.T)) {
            float4 o;
            o.x = r[4 * i] + b2;
            o.y = r[4 * i + 1] + b2;
            o.z = r[4 * i + 2] + b2;
            o.w = r[4 * i + 3] + b2;
            *reinterpret_cast<float4*>(yr + ch * 32 + 4 * i) = o;  // the last group may spill into the row padding
          }
        }
      }
      tc_fence_before();
      prefill(it + 2);  // same accumulator stage
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, PASS == 1 ? 256u : 512u);
}

}  // namespace
}  // namespace sty

static int sty_convnext_fused_eligible(const float* x, int64_t x_bs, int64_t x_cs, const float* y, int64_t y_bs,
                                       int64_t y_cs, int C, int J, int T) {
  return C == sty::kC && J == sty::kJ && T >= 512 && x != y && sty::tma_layout_ok(x, x_bs, x_cs) &&
         sty::tma_layout_ok(y, y_bs, y_cs) && x_cs >= ((T + 3) & ~3) && y_cs >= ((T + 3) & ~3);
}

extern "C" int sty_convnext_fused_fwd(const float* x, int64_t x_bs, int64_t x_cs, float* y, int64_t y_bs, int64_t y_cs,
                                      const float* dw_w, const float* dw_b, const float* gb, int64_t gb_bs, float eps,
                                      const void* w1_split, const float* b1, const float* alpha,
                                      const float* grn_gamma, const void* w2_split, const float* b2, float* sumsq,
                                      float* gs, int B, int C, int J, int T, sty_stream_t stream) {
  using namespace sty;
  STY_REQUIRE(x && y && dw_w && dw_b && gb && w1_split && b1 && alpha && grn_gamma && w2_split && b2 && sumsq && gs,
              "convnext_fused: null pointer");
  STY_REQUIRE(B > 0 && sty_convnext_fused_eligible(x, x_bs, x_cs, y, y_bs, y_cs, C, J, T),
              "convnext_fused: needs C=32, 4C=128, T>=512, out != in, 16-byte aligned rows padded to 4 steps");
  cudaStream_t st = as_stream(stream);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  STY_REQUIRE(make_tmap_bct(&tmap, x, B, C, T, x_bs, x_cs, kRawW, kC), "convnext_fused: tensor map encoding failed");
  Args a;
  a.x = x; a.x_bs = x_bs; a.x_cs = x_cs;
  a.y = y; a.y_bs = y_bs; a.y_cs = y_cs;
  a.dw_w = dw_w; a.dw_b = dw_b; a.gb = gb; a.gb_bs = gb_bs; a.eps = eps;
  a.w1s = reinterpret_cast<const uint4*>(w1_split);
  a.w2s = reinterpret_cast<const uint4*>(w2_split);
  a.b1 = b1; a.alpha = alpha; a.b2 = b2;
  a.sumsq = sumsq; a.gs = gs;
  a.B = B; a.T = T; a.tiles_per_b = cdiv(T, kMT);
  static int sms = 0;
  if (sms <= 0) {
    sms = sty_device_sm_count();
    if (sms <= 0) sms = 148;
  }
  const int64_t n_tiles = (int64_t)B * a.tiles_per_b;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (cudaMemsetAsync(sumsq, 0, (size_t)B * J * sizeof(float), st) != cudaSuccess) {
    set_error("convnext_fused: memset failed");
    return STY_ERR_CUDA;
  }
  cudaFuncSetAttribute(convnext_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem1);
  cudaFuncSetAttribute(convnext_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2);
  convnext_fused_kernel<1><<<grid, kThreads1, kSmem1, st>>>(a, tmap);
  STY_CHECK_LAUNCH("convnext_fused pass 1");
  const int rc = sty_grn_scale_fwd(sumsq, grn_gamma, gs, B, J, stream);
  if (rc != STY_OK) return rc;
  convnext_fused_kernel<2><<<grid, kThreads2, kSmem2, st>>>(a, tmap);
  STY_CHECK_LAUNCH("convnext_fused pass 2");
  return STY_OK;
}
