// Conv1d on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM),
// im2col-free: the input tile is staged ONCE in shared memory in the UMMA
// canonical K-major / no-swizzle layout with rows (time steps) uniformly 16 B
// apart, so the K taps of the convolution are the SAME staged tile addressed
// through shared-memory descriptors whose start address is advanced by
// tap*dilation rows.  Per tile:  D[t, co] += sum_tap  X[t + tap*dil, ci] * W_tap[ci, co].
//
// Precision: fp32 operands are split into bf16 (hi, lo) pairs while staging and
// three MMAs (hi*hi + lo*hi + hi*lo) are accumulated in fp32 ("bf16x3"), which
// keeps ~16 mantissa bits per operand — plain bf16/tf32 MMA breaks the 1e-3
// end-to-end parity bound.
//
// The fused prologue (mask, AdaIN/GRN affine, LeakyReLU/Snake) runs while
// staging; the epilogue (bias, activation, mask, scale, residual, GRN sum of
// squares, pixel-shuffle store) runs on the accumulator rows read back with
// tcgen05.ld.  Semantics identical to the SIMT kernel in conv1d.cu.
//
// Execution: persistent, warp-specialised CTAs (one per SM).  Warps 0-7 stage
// input tiles into a ring of shared-memory stages (up to 32 global loads in
// flight per thread), warp 8 issues the MMAs (one elected thread) into one of
// two TMEM accumulator stages, warps 9-16 drain the other accumulator stage
// (epilogue) — the three roles hand over through mbarriers, so staging of tile
// i+1, MMA of tile i and the epilogue of tile i-1 overlap.  Weights stay resident in shared memory when they fit, otherwise
// they are streamed per input-channel chunk together with the input rows.
#include <stdlib.h>
#include <string.h>

#include "tma.cuh"

namespace sty {

// -------------------------------------------------------------------- kernel
struct UmmaPlan {
  int ci_chunk;    // input channels per staged chunk (multiple of 16)
  int n_chunks;    // ceil(CI / ci_chunk)
  int rows;        // staged time steps per tile: 128 + (K-1)*dil
  int NT;          // output channels per CTA (multiple of 16, <= 256)
  int n_stages;    // shared-memory ring depth
  int resident;    // 1: all weights of this CTA's NT channels stay in shared memory
  int acc_cols;    // TMEM columns per accumulator stage (power of two >= NT (2*NT when dual), >= 32)
  int dual;        // 1: NT <= 64 — hi*hi and hi*lo come from ONE MMA against [W_hi | W_lo] (N = 2*NT): 2 MMAs, not 3
  int stage_u4;    // uint4 per stage
  int wres_u4;     // uint4 of the resident weight block (0 when streaming)
  int tiles_per_b; // ceil(T / 128)
  int prm_floats;  // floats of the per-channel parameter block (prologue + epilogue)
  int dw_floats;   // IN_MODE 4: two raw input tiles [2][CI][134+1]
  int tma;         // 1: raw fp32 input tiles arrive by TMA (cp.async.bulk.tensor) into a ring of raw stages
  int raw_rp;      // TMA: floats per channel row of a raw stage (rows rounded up to a multiple of 4)
  int raw_stages;  // TMA: ring depth
  int raw_u4;      // TMA: uint4 per raw stage
  int raw_shift;   // TMA: the box starts at t0 - pad - raw_shift, a multiple of 4 steps (the hardware wants the
                   // first element of a box on a 16-byte boundary, measured: tools/scratch/tma_test.cu)
};

constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kEpilogueWarp0 = kMmaWarp + 1;
constexpr int kEpilogueWarps = 8;  // 4 lane quadrants x kColParts column parts (12 or 16 warps cap the registers at 80 / 72 and spill in the prologue producers: measured slower)
constexpr int kColParts = kEpilogueWarps / 4;
constexpr int kThreads = (kEpilogueWarp0 + kEpilogueWarps) * 32;
constexpr int kMaxStages = 4;
constexpr int kMaxRaw = 4;
constexpr int kLoaderWarp = kEpilogueWarp0 + kEpilogueWarps;  // TMA variants only: one more warp
constexpr int kThreadsTma = kThreads + 32;
constexpr int kItemBatch = 4;  // (row, 8-channel) items staged per thread per batch: 32 loads in flight

// column sums over the 32 lanes of 16 per-lane values: lanes l and l+16 return sum_lanes v[l & 15]
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// MMAs of one staged chunk (all taps, all 16-channel K blocks), issued by the elected thread.  Only the first MMA
// carries the run-time accumulate flag; everything else is branch-free so that the issuing thread — the kernel's
// critical path for K = 11 / 21 (one thread feeds the tensor pipe) — executes a handful of uniform instructions
// per MMA.  DUAL: hi*hi and hi*lo come from ONE MMA against [W_hi | W_lo] (N = 2*NT).
template <bool DUAL>
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t a0, uint32_t a_hi, uint32_t b0, uint32_t b_hi,
                                            uint32_t idesc, uint32_t idesc2, int K, uint32_t dil, int kblocks,
                                            uint32_t a_kstep, uint32_t b_kstep, uint32_t a_lo_off, uint32_t b_lo_off,
                                            uint32_t b_tstep, uint32_t acc0) {
  auto block = [&](uint32_t ak, uint32_t bk, uint32_t acc) {
    if constexpr (DUAL) {
      umma_bf16_w(d_tmem, ak, a_hi, bk, b_hi, idesc2, acc);            // hi * [hi | lo]
      umma_bf16_w(d_tmem, ak + a_lo_off, a_hi, bk, b_hi, idesc, 1u);   // lo * hi -> cols [0,NT)
    } else {
      umma_bf16_w(d_tmem, ak, a_hi, bk, b_hi, idesc, acc);             // hi * hi
      umma_bf16_w(d_tmem, ak + a_lo_off, a_hi, bk, b_hi, idesc, 1u);   // lo * hi
      umma_bf16_w(d_tmem, ak, a_hi, bk + b_lo_off, b_hi, idesc, 1u);   // hi * lo
    }
  };
  block(a0, b0, acc0);
  {
    uint32_t ak = a0 + a_kstep, bk = b0 + b_kstep;
#pragma unroll 2
    for (int kb = 1; kb < kblocks; ++kb, ak += a_kstep, bk += b_kstep) block(ak, bk, 1u);
  }
  uint32_t a_t = a0 + dil, b_t = b0 + b_tstep;
#pragma unroll 1
  for (int tap = 1; tap < K; ++tap, a_t += dil, b_t += b_tstep) {
    uint32_t ak = a_t, bk = b_t;
#pragma unroll 2
    for (int kb = 0; kb < kblocks; ++kb, ak += a_kstep, bk += b_kstep) block(ak, bk, 1u);
  }
}

// IN_MODE : 0 no prologue | 1 mask/affine | 2 mask/affine + LeakyReLU(0.2) | 3 mask/affine + Snake | 5 mask/affine + any
//           other activation (act_apply: GELU, ReLU, Swish)
//           4 ConvNeXt front: depthwise k7 conv + LayerNorm over channels + adaptive affine (K=1 only)
// OUT_MODE: 0 none | 1 Snake | 2 ReLU | 3 Swish          (compile-time: keeps each role's loop small
// enough for the instruction cache — three roles run different code on one SM)
// EPI     : 0 general epilogue | 1 no residual / mask / output scale / pixel shuffle | 2 residual with
//           res_scale 1, no mask / scale / shuffle.  Compile-time because the three roles share one
//           instruction cache: dropping the unused epilogue paths took the fused-front kernel from 0.267 to
//           0.242 ms.  Only the hot (IN_MODE, OUT_MODE, EPI) combinations are instantiated (pick_kernel).
// TMA     : the raw fp32 input tile [ci_chunk][rows] is fetched by one cp.async.bulk.tensor per stage issued by a
//           loader warp (zero fill outside [0, T)), several stages ahead; the producer warps then convert FROM
//           SHARED MEMORY (prologue + bf16 hi/lo split).  No register is held across a global-memory latency,
//           and the bytes in flight per SM are raw_stages x stage size (54-70 KB) instead of 32 KB.
template <int IN_MODE, int OUT_MODE, int EPI = 0, bool TMA = false>
__global__ void __launch_bounds__(TMA ? kThreadsTma : kThreads, 1)
conv1d_umma_kernel(const sty_conv1d_args p, const UmmaPlan pl, const __grid_constant__ CUtensorMap tmap) {
  static_assert(!(TMA && IN_MODE == 4), "the fused ConvNeXt front has its own kernel for TMA");
  constexpr bool PRO = (IN_MODE >= 1 && IN_MODE <= 3) || IN_MODE == 5;
  constexpr int MT = 128;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int K = p.K, dil = p.dil, CO = p.CO, CI = p.CI, NT = pl.NT, rows = pl.rows;
  const int NS = pl.n_stages;
  // compile-time off in the fused-front kernels (NT >= 128 there): their loops are i-cache sensitive
  const bool dual = IN_MODE != 4 && pl.dual != 0;
  // LEAN epilogue: bias + activation + sums only (always for the fused ConvNeXt front: the host only selects
  // IN_MODE 4 when there is no residual, mask, output scale, pixel shuffle or plain sum); RES1: + residual
  constexpr bool LEAN = IN_MODE == 4 || EPI == 1;
  constexpr bool RES1 = EPI == 2;
  const int c8c = pl.ci_chunk >> 3;  // 16-byte K chunks per staged chunk
  const int co0 = blockIdx.y * NT;
  uint4* Wres = reinterpret_cast<uint4*>(smem_raw);       // resident: [K][2][CI/8][NT]
  uint4* stage0 = Wres + pl.wres_u4;                       // stage: X [2][c8c][rows] (+ W [K][2][c8c][NT])
  float* prm = reinterpret_cast<float*>(stage0 + (size_t)NS * pl.stage_u4);
  float* pro_s = prm;                // [4][CI]: scale, shift, alpha, 1/alpha of the current batch element
  float* epi_s = prm + (IN_MODE == 4 ? 10 : 4) * CI;  // [3][NT]: bias, alpha, 1/alpha
  float* dw_s = prm + pl.prm_floats;  // IN_MODE 4 scratch
  // TMA raw ring: 128-byte aligned (all preceding blocks are multiples of 16 B; round up)
  uint8_t* raw_base = reinterpret_cast<uint8_t*>(dw_s + pl.dw_floats);
  raw_base += (128u - (smem_u32(raw_base) & 127u)) & 127u;
  float* raw0 = reinterpret_cast<float*>(raw_base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_base + (TMA ? (size_t)pl.raw_stages * pl.raw_u4 * 16 : 0));
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + kMaxStages;
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* raw_full = acc_empty + 2;
  uint64_t* raw_empty = raw_full + kMaxRaw;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kMaxRaw);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint4* __restrict__ wsplit = reinterpret_cast<const uint4*>(p.w_split);

  if (warp == kMmaWarp) tmem_alloc(tmem_slot, (uint32_t)(2 * pl.acc_cols));
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&x_full[i], IN_MODE == 4 ? 128 : kProducerThreads);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpilogueWarps * 32);
    }
    if (TMA) {
      for (int i = 0; i < pl.raw_stages; ++i) {
        mbar_init(&raw_full[i], 1);
        mbar_init(&raw_empty[i], kProducerThreads);
      }
    }
    fence_barrier_init();
  }
  for (int i = tid; i < NT; i += (TMA ? kThreadsTma : kThreads)) {
    const float al = p.out_alpha ? p.out_alpha[co0 + i] : 1.f;
    epi_s[i] = p.bias ? p.bias[co0 + i] : 0.f;
    epi_s[NT + i] = al;
    epi_s[2 * NT + i] = 1.0f / al;
  }
  if (pl.resident) {
    // all weights of this CTA's output channels, staged once: [K*2 blocks][CI/8][NT]
    // (dual: [K][CI/8][hi | lo][NT], so that one descriptor spans the hi and the lo rows of a K chunk)
    const int total = K * 2 * (CI >> 3) * NT;
    const int c8n = CI >> 3;
    for (int idx = tid; idx < total; idx += (TMA ? kThreadsTma : kThreads)) {
      const int row = idx / NT, n = idx - row * NT;  // row = blk*(CI/8) + c8, blk = tap*2 + split
      int dst = idx;
      if (dual) {
        const int blk = row / c8n, c8 = row - blk * c8n;
        dst = (((blk >> 1) * c8n + c8) * 2 + (blk & 1)) * NT + n;
      }
      Wres[dst] = wsplit[(int64_t)row * CO + co0 + n];
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = (int64_t)p.B * pl.tiles_per_b;
  const int tile_begin = (int)((n_tiles * blockIdx.x) / gridDim.x);
  const int tile_end = (int)((n_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp < kProducerWarps) {
    // =========================== producers: stage input rows (and streamed weights)
    uint32_t it = 0;
    int cur_b = -1;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int b = tile / pl.tiles_per_b;
      const int t0 = (tile - b * pl.tiles_per_b) * MT;
      const float* __restrict__ xb = p.x + (int64_t)b * p.x_bs;
      const float* __restrict__ in_mask = (PRO && p.in_mask) ? p.in_mask + (int64_t)b * p.T : nullptr;
      if (PRO && b != cur_b) {  // per-channel prologue parameters of this batch element -> smem
        asm volatile("bar.sync 2, %0;" ::"n"(kProducerThreads) : "memory");
        for (int c = tid; c < CI; c += kProducerThreads) {
          const float al = p.in_alpha ? p.in_alpha[c] : 1.f;
          pro_s[c] = p.in_scale ? p.in_scale[(int64_t)b * CI + c] : 1.f;
          pro_s[CI + c] = p.in_shift ? p.in_shift[(int64_t)b * CI + c] : 0.f;
          pro_s[2 * CI + c] = al;
          pro_s[3 * CI + c] = 1.0f / al;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(kProducerThreads) : "memory");
        cur_b = b;
      }
      if constexpr (IN_MODE == 4) {
        // ---- ConvNeXt front (conv_next.py:82-84): y = (1+g) * LN_C(dwconv7(x) + b) + beta.
        // Two producer groups of 128 threads take alternate tiles (two tiles in flight); a thread
        // owns one time step with ALL channels, so LayerNorm needs no cross-thread exchange.
        constexpr int RP = 128 + 6 + 1;  // raw row pitch
        const int grp = tid >> 7, r = tid & 127;
        float* raw = dw_s + grp * (CI * RP);  // [CI][RP] raw fp32 tile of this group
        float* dwp = pro_s;                    // [CI][8] taps+bias, then gamma[CI], beta[CI]
        if (b != cur_b) {
          asm volatile("bar.sync 2, %0;" ::"n"(kProducerThreads) : "memory");
          for (int i = tid; i < CI * 8; i += kProducerThreads) {
            const int c = i >> 3, k = i & 7;
            dwp[i] = k < 7 ? p.dw_w[c * 7 + k] : p.dw_b[c];
          }
          for (int c = tid; c < 2 * CI; c += kProducerThreads)
            dwp[CI * 8 + c] = p.dw_gb[(int64_t)b * p.dw_gb_bs + c];
          asm volatile("bar.sync 2, %0;" ::"n"(kProducerThreads) : "memory");
          cur_b = b;
        }
        if ((it & 1) == (uint32_t)grp) {
          const int s = it % NS;
          mbar_wait_sleep(&x_empty[s], ((it / NS) & 1) ^ 1);
          uint4* Xs = stage0 + (size_t)s * pl.stage_u4;
          // group-local barrier (ids 3, 4): previous tile's readers of `raw` are done
          if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 4, 128;" ::: "memory");
          {
            const int ta = t0 - 3 + r, tb = ta + 128;
            const bool oka = ta >= 0 && ta < p.T, okb = (r < 6) && tb >= 0 && tb < p.T;
            // 32 independent loads in flight per thread before the first store (latency-bound otherwise)
            const int64_t cs = p.x_cs;
            for (int cb = 0; cb < CI; cb += 32) {
              float va[32];
              if (cb + 32 <= CI) {  // whole batch of channels exists: running pointer, one predicate
                const float* __restrict__ sp = xb + (int64_t)cb * cs + ta;
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                  va[u] = oka ? *sp : 0.f;
                  sp += cs;
                }
#pragma unroll
                for (int u = 0; u < 32; ++u) raw[(cb + u) * RP + r] = va[u];
              } else {
#pragma unroll
                for (int u = 0; u < 32; ++u)
                  va[u] = (oka && cb + u < CI) ? xb[(int64_t)(cb + u) * cs + ta] : 0.f;
#pragma unroll
                for (int u = 0; u < 32; ++u)
                  if (cb + u < CI) raw[(cb + u) * RP + r] = va[u];
              }
            }
            if (r < 6) {
              for (int cb = 0; cb < CI; cb += 16) {
                float vb[16];
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  vb[u] = (okb && cb + u < CI) ? xb[(int64_t)(cb + u) * p.x_cs + tb] : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u)
                  if (cb + u < CI) raw[(cb + u) * RP + 128 + r] = vb[u];
              }
            }
          }
          if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
          else asm volatile("bar.sync 4, 128;" ::: "memory");
          auto dwc = [&](int c) {
            const float* wr = dwp + c * 8;
            const float* xr = raw + c * RP + r;
            float a = wr[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) a = fmaf(wr[k], xr[k], a);
            return a;
          };
          const bool ok = (t0 + r) < p.T;
          const float invC = 1.0f / (float)CI;
          auto emit = [&](int c8, const float* y8) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a0 = ok ? y8[2 * j] : 0.f, a1 = ok ? y8[2 * j + 1] : 0.f;
              h[j] = pack_bf16(a0, a1);
              l[j] = pack_bf16(a0 - __uint_as_float(h[j] << 16), a1 - __uint_as_float(h[j] & 0xffff0000u));
            }
            Xs[(0 * c8c + c8) * rows + r] = make_uint4(h[0], h[1], h[2], h[3]);
            Xs[(1 * c8c + c8) * rows + r] = make_uint4(l[0], l[1], l[2], l[3]);
          };
          if (CI == 32) {  // whole column in registers
            float d[32];
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              d[c] = dwc(c);
              sum += d[c];
            }
            const float mean = sum * invC;
            float q = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float e = d[c] - mean;
              q = fmaf(e, e, q);
            }
            const float rstd = 1.0f / sqrtf(q * invC + p.dw_eps);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              float y8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = c8 * 8 + j;
                y8[j] = (1.0f + dwp[CI * 8 + c]) * ((d[c] - mean) * rstd) + dwp[CI * 9 + c];
              }
              emit(c8, y8);
            }
          } else {  // recompute the cheap depthwise conv per pass instead of holding CI values
            float sum = 0.f;
            for (int c = 0; c < CI; ++c) sum += dwc(c);
            const float mean = sum * invC;
            float q = 0.f;
            for (int c = 0; c < CI; ++c) {
              const float e = dwc(c) - mean;
              q = fmaf(e, e, q);
            }
            const float rstd = 1.0f / sqrtf(q * invC + p.dw_eps);
            for (int c8 = 0; c8 < (CI >> 3); ++c8) {
              float y8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = c8 * 8 + j;
                y8[j] = (1.0f + dwp[CI * 8 + c]) * ((dwc(c) - mean) * rstd) + dwp[CI * 9 + c];
              }
              emit(c8, y8);
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(&x_full[s]);
        }
        ++it;
      } else
      for (int ch = 0; ch < pl.n_chunks; ++ch, ++it) {
        const int c0 = ch * pl.ci_chunk;
        const int cc8 = min(pl.ci_chunk, CI - c0) >> 3;
        const int s = it % NS;
        mbar_wait_sleep(&x_empty[s], ((it / NS) & 1) ^ 1);
        uint4* Xs = stage0 + (size_t)s * pl.stage_u4;
        if (!pl.resident) {
          uint4* Ws = Xs + 2 * c8c * rows;  // [K*2 blocks][c8c][NT]
          const int blk_elems = cc8 * NT;
          const int total = K * 2 * blk_elems;
#pragma unroll 4
          for (int idx = tid; idx < total; idx += kProducerThreads) {
            const int blk = idx / blk_elems, within = idx - blk * blk_elems;
            const int c8l = within / NT, n = within - c8l * NT;
            const int dst = dual ? (((blk >> 1) * c8c + c8l) * 2 + (blk & 1)) * NT + n
                                    : blk * (c8c * NT) + within;
            Ws[dst] = wsplit[((int64_t)blk * (CI >> 3) + (c0 >> 3) + c8l) * CO + co0 + n];
          }
        }
        // items: (c8, row) pairs, row fastest; thread takes items tid, tid+256, ...
        const int64_t x_cs = p.x_cs;
        const int n_items = cc8 * rows;
        int row = tid, c8 = 0;
        while (row >= rows && c8 < cc8) { row -= rows; ++c8; }
        const int rs = TMA ? (int)(it % (uint32_t)pl.raw_stages) : 0;
        const float* raw = raw0 + (size_t)rs * pl.raw_u4 * 4;  // [ci_chunk][raw_rp] fp32, row 0 = t0 - pad
        if constexpr (TMA) mbar_wait(&raw_full[rs], (it / (uint32_t)pl.raw_stages) & 1);
        for (int i0 = tid; i0 < n_items; i0 += kProducerThreads * kItemBatch) {
          float v[kItemBatch][8];
          int irow[kItemBatch], ic8[kItemBatch];
#pragma unroll
          for (int u = 0; u < kItemBatch; ++u) {
            irow[u] = row;
            ic8[u] = c8;
            if constexpr (TMA) {
              // shared-memory reads: lanes = consecutive rows -> conflict-free; out-of-range steps are zeros
              const float* src = raw + (c8 < cc8 ? c8 * 8 : 0) * pl.raw_rp + row + pl.raw_shift;
#pragma unroll
              for (int j = 0; j < 8; ++j) v[u][j] = src[j * pl.raw_rp];
            } else {
              const int t = t0 - p.pad + row;
              const bool ok = (c8 < cc8) && (t >= 0) && (t < p.T);
              const float* __restrict__ src = xb + (int64_t)(c0 + c8 * 8) * x_cs + t;
#pragma unroll
              for (int j = 0; j < 8; ++j) {  // running pointer: 2 integer instructions per load instead of ~5
                v[u][j] = ok ? *src : 0.f;
                src += x_cs;
              }
            }
            row += kProducerThreads;
            while (row >= rows && c8 < cc8) { row -= rows; ++c8; }
          }
#pragma unroll
          for (int u = 0; u < kItemBatch; ++u) {
            if (ic8[u] < cc8) {
              const int t = t0 - p.pad + irow[u];
              const bool ok = (t >= 0) && (t < p.T);
              if constexpr (PRO) {
                const int cbase = c0 + ic8[u] * 8;
                const float m = (ok && in_mask) ? in_mask[t] : 1.f;
                // per-channel parameters as 128-bit shared loads (cbase is a multiple of 8, CI of 16)
                float sc[8], sh[8], al[8], ial[8];
                *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(pro_s + cbase);
                *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(pro_s + cbase + 4);
                *reinterpret_cast<float4*>(sh) = *reinterpret_cast<const float4*>(pro_s + CI + cbase);
                *reinterpret_cast<float4*>(sh + 4) = *reinterpret_cast<const float4*>(pro_s + CI + cbase + 4);
                if constexpr (IN_MODE == 3) {
                  *reinterpret_cast<float4*>(al) = *reinterpret_cast<const float4*>(pro_s + 2 * CI + cbase);
                  *reinterpret_cast<float4*>(al + 4) = *reinterpret_cast<const float4*>(pro_s + 2 * CI + cbase + 4);
                  *reinterpret_cast<float4*>(ial) = *reinterpret_cast<const float4*>(pro_s + 3 * CI + cbase);
                  *reinterpret_cast<float4*>(ial + 4) = *reinterpret_cast<const float4*>(pro_s + 3 * CI + cbase + 4);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float w = fmaf(v[u][j] * m, sc[j], sh[j]);
                  if constexpr (IN_MODE == 3) {
                    w = fmaf(ial[j], sin_sq(al[j] * w), w);
                  } else if constexpr (IN_MODE == 2) {
                    w = w > 0.f ? w : 0.2f * w;
                  } else if constexpr (IN_MODE == 5) {
                    w = act_apply(w, p.in_act);
                  }
                  v[u][j] = (ok && m >= 0.f) ? w : 0.f;  // negative mask value: zero AFTER the prologue
                }
              }
              uint32_t h[4], l[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float a0 = v[u][2 * j], a1 = v[u][2 * j + 1];
                h[j] = pack_bf16(a0, a1);
                const float l0 = a0 - __uint_as_float(h[j] << 16);
                const float l1 = a1 - __uint_as_float(h[j] & 0xffff0000u);
                l[j] = pack_bf16(l0, l1);
              }
              Xs[(0 * c8c + ic8[u]) * rows + irow[u]] = make_uint4(h[0], h[1], h[2], h[3]);
              Xs[(1 * c8c + ic8[u]) * rows + irow[u]] = make_uint4(l[0], l[1], l[2], l[3]);
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core proxy
        mbar_arrive(&x_full[s]);
        if constexpr (TMA) mbar_arrive(&raw_empty[rs]);  // raw stage read: the loader may refill it
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (one elected lane)
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = NT, M = 128
    const uint32_t idesc =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
    const int wc8 = pl.resident ? (CI >> 3) : c8c;  // K-chunk rows per (tap, split) weight block
    uint32_t it = 0, j = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++j) {
      const uint32_t a = j & 1;
      mbar_wait(&acc_empty[a], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (uint32_t)pl.acc_cols;
      for (int ch = 0; ch < pl.n_chunks; ++ch, ++it) {
        const int c0 = ch * pl.ci_chunk;
        const int cc8 = min(pl.ci_chunk, CI - c0) >> 3;
        const int s = it % NS;
        mbar_wait(&x_full[s], (it / NS) & 1);
        tc_fence_after();
        if (elect_one()) {
          // single issuing thread; only the low descriptor word (start address) changes per MMA
          const uint4* Xs = stage0 + (size_t)s * pl.stage_u4;
          const uint32_t ws_addr = pl.resident ? smem_u32(Wres) : smem_u32(Xs + 2 * c8c * rows);
          const uint64_t a_d = make_desc(smem_u32(Xs), (uint32_t)rows, 8u);
          const uint64_t b_d = make_desc(ws_addr, (uint32_t)(dual ? 2 * NT : NT), 8u);
          const uint32_t a_hi32 = (uint32_t)(a_d >> 32), b_hi32 = (uint32_t)(b_d >> 32);
          uint32_t a_t = (uint32_t)a_d, b_t = (uint32_t)b_d;  // low words: (tap 0, kb 0, hi split)
          // resident weights with a chunked input (TMA plans): this chunk's K rows of the weight block
          if (pl.resident) b_t += (uint32_t)((c0 >> 3) * (dual ? 2 * NT : NT));
          const uint32_t a_lo_off = (uint32_t)(c8c * rows);    // 16-byte units to the lo-split copy
          const uint32_t b_lo_off = (uint32_t)(wc8 * NT);
          const int kblocks = cc8 >> 1;                         // MMA K = 16 bf16 = two 16-byte chunks
          const uint32_t a_kstep = 2 * rows, b_kstep = dual ? 4 * NT : 2 * NT;
          const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * NT) >> 3) << 17);  // N = 2*NT
          const uint32_t b_tstep = 2 * wc8 * NT;                // per tap (hi and lo blocks)
          if (dual)
            issue_chunk<true>(d_tmem, a_t, a_hi32, b_t, b_hi32, idesc, idesc2, K, (uint32_t)dil, kblocks, a_kstep,
                              b_kstep, a_lo_off, b_lo_off, b_tstep, ch > 0 ? 1u : 0u);
          else
            issue_chunk<false>(d_tmem, a_t, a_hi32, b_t, b_hi32, idesc, idesc2, K, (uint32_t)dil, kblocks, a_kstep,
                               b_kstep, a_lo_off, b_lo_off, b_tstep, ch > 0 ? 1u : 0u);
          umma_commit(&x_empty[s]);                               // stage free once read
          if (ch == pl.n_chunks - 1) umma_commit(&acc_full[a]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (TMA && warp == kLoaderWarp) {
    // =========================== TMA loader: one lane keeps raw_stages tiles in flight
    if (elect_one()) {
      prefetch_tensormap(&tmap);
      const uint32_t NR = (uint32_t)pl.raw_stages;
      const uint32_t bytes = (uint32_t)pl.ci_chunk * (uint32_t)pl.raw_rp * 4u;
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int b = tile / pl.tiles_per_b;
        const int t0 = (tile - b * pl.tiles_per_b) * MT;
        for (int ch = 0; ch < pl.n_chunks; ++ch, ++it) {
          const uint32_t r = it % NR;
          mbar_wait_sleep(&raw_empty[r], ((it / NR) & 1) ^ 1);
          mbar_arrive_expect_tx(&raw_full[r], bytes);
          tma_load_3d(raw0 + (size_t)r * pl.raw_u4 * 4, &tmap, &raw_full[r], t0 - p.pad - pl.raw_shift,
                      ch * pl.ci_chunk, b);
        }
      }
    }
  } else {
    // =========================== epilogue: thread owns accumulator row (TMEM lane) q*32+lane
    // and the 16-column chunks c with (c % kColParts) == half
    const int ew = warp - kEpilogueWarp0;
    const int q = warp & 3;
    const int half = ew >> 2;
    const int s = p.shuffle > 1 ? p.shuffle : 1;
    const int n_chunks16 = NT >> 4;
    float ssq_acc[8];  // column (16*c + (lane & 15)) sums of this thread's chunks c = half, half+kColParts, ...
#pragma unroll
    for (int i = 0; i < 8; ++i) ssq_acc[i] = 0.f;
    float sum_acc[2] = {0.f, 0.f};  // same for out_sum (NT <= 64: at most two chunks per thread)
    uint32_t j = 0;
    int ssq_b = -1;
    auto flush_ssq = [&](int bb) {
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = half + kColParts * i;
          if (p.out_sumsq && c < n_chunks16)
            atomicAdd(p.out_sumsq + (int64_t)bb * CO + co0 + c * 16 + lane, ssq_acc[i]);
          ssq_acc[i] = 0.f;
        }
        if (p.out_sum) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c = half + kColParts * i;
            if (c < n_chunks16) atomicAdd(p.out_sum + (int64_t)bb * CO + co0 + c * 16 + lane, sum_acc[i]);
            sum_acc[i] = 0.f;
          }
        }
      }
    };
    for (int tile = tile_begin; tile < tile_end; ++tile, ++j) {
      const int b = tile / pl.tiles_per_b;
      const int t0 = (tile - b * pl.tiles_per_b) * MT;
      const uint32_t a = j & 1;
      if ((p.out_sumsq || p.out_sum) && ssq_b != b) {
        if (ssq_b >= 0) flush_ssq(ssq_b);
        ssq_b = b;
      }
      mbar_wait_sleep(&acc_full[a], (j >> 1) & 1);
      tc_fence_after();
      const int t = t0 + q * 32 + lane;
      const bool t_ok = t < p.T;
      const float* __restrict__ out_mask = p.out_mask ? p.out_mask + (int64_t)b * p.T : nullptr;
      float* __restrict__ yb = p.y + (int64_t)b * p.y_bs;
      const float* __restrict__ rb = p.res ? p.res + (int64_t)b * p.r_bs : nullptr;
      const float om = ((out_mask && t_ok) ? out_mask[t] : 1.f) * p.out_scale;
      const uint32_t acc_addr = tmem_base + a * (uint32_t)pl.acc_cols + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int ci = 0; ci < 8; ++ci) {
        const int c = half + kColParts * ci;
        if (c >= n_chunks16) break;
        const int n0 = c * 16;
        float r[16], rv[16];
        // residual rows first: all 16 loads in flight before any store (res may alias y)
        if (!LEAN && rb && t_ok && (RES1 || s == 1)) {
          const float* __restrict__ rp = rb + (int64_t)(co0 + n0) * p.r_cs + t;  // running pointer, see producers
          const int64_t r_cs = p.r_cs;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            rv[jj] = *rp;
            rp += r_cs;
          }
        } else if (!LEAN && rb && t_ok) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int co = co0 + n0 + jj;
            if (s == 1) {
              rv[jj] = rb[(int64_t)co * p.r_cs + t];
            } else {
              const int c_out = co / s, r_out = co - c_out * s;
              rv[jj] = rb[(int64_t)c_out * p.r_cs + (int64_t)t * s + r_out];
            }
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) rv[jj] = 0.f;
        }
        tmem_ld16(acc_addr + (uint32_t)n0, r);
        if (dual) {  // columns [NT, 2*NT) hold the hi*lo partial products
          float r2[16];
          tmem_ld16(acc_addr + (uint32_t)(NT + n0), r2);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) r[jj] += r2[jj];
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const int n = n0 + jj;
          float v = r[jj] + epi_s[n];
          if constexpr (OUT_MODE == 1) {
            v = fmaf(epi_s[2 * NT + n], sin_sq(epi_s[NT + n] * v), v);
          } else if constexpr (OUT_MODE == 2) {
            v = fmaxf(v, 0.f);
          } else if constexpr (OUT_MODE == 3) {
            v = v / (1.0f + __expf(-v));
          }
          if constexpr (RES1) v += rv[jj];
          else if constexpr (!LEAN) v = fmaf(p.res_scale, rv[jj], v * om);
          if (!t_ok) v = 0.f;
          r[jj] = v;
        }
        if (t_ok) {
          if (LEAN || RES1 || s == 1) {
            float* __restrict__ yp = yb + (int64_t)(co0 + n0) * p.y_cs + t;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) yp[(int64_t)jj * p.y_cs] = r[jj];
          } else {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int co = co0 + n0 + jj;
              const int c_out = co / s, r_out = co - c_out * s;
              yb[(int64_t)c_out * p.y_cs + (int64_t)t * s + r_out] = r[jj];
            }
          }
        }
        if (IN_MODE != 4 && p.out_sum) {
          float cp[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) cp[jj] = r[jj];
          const float colsum = warp_colsum16(cp, lane);
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (i == ci) sum_acc[i] += colsum;
        }
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) r[jj] *= r[jj];
        if (p.out_sumsq) {
          const float colsum = warp_colsum16(r, lane);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == ci) ssq_acc[i] += colsum;
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[a]);
    }
    if ((p.out_sumsq || p.out_sum) && ssq_b >= 0) flush_ssq(ssq_b);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)(2 * pl.acc_cols));
}

// ------------------------------------------------------------------ host side
static const size_t kSmemBudget = 226 * 1024;
static const size_t kSmemBars = 256;  // barriers + tmem slot

static bool make_plan(const sty_conv1d_args& a, UmmaPlan& pl) {
  if (a.CI % 16 != 0 || a.CO % 16 != 0 || a.CO < 16 || a.CI > 4096) return false;
  pl.rows = 128 + (a.K - 1) * a.dil;
  if (pl.rows >= 16384) return false;
  pl.tiles_per_b = cdiv(a.T, 128);
  // output-channel tile candidates: CO itself if <= 128, else multiple-of-16 divisors, largest first
  // (each epilogue thread keeps GRN partial sums for at most 8 of its 16-column chunks => NT <= 256)
  int nt_first = a.CO <= 256 ? a.CO : 256;
  {
    // few tiles (text-rate convs: 16 x 258 steps = 33 tiles): prefer narrower output-channel tiles while twice
    // as many CTAs still fit on the SMs, instead of leaving three quarters of the GPU idle
    const int64_t n_tiles = (int64_t)a.B * pl.tiles_per_b;
    while (nt_first >= 32) {
      while (nt_first >= 16 && a.CO % nt_first != 0) nt_first -= 16;
      if (nt_first < 32 || n_tiles * (a.CO / nt_first) * 2 > 148) break;
      int d = nt_first - 16;
      while (d >= 16 && a.CO % d != 0) d -= 16;
      if (d < 16) break;
      nt_first = d;
    }
  }
  for (int nt = nt_first; nt >= 16; nt -= 16) {
    if (a.CO % nt != 0) continue;
    if (a.CO / nt > 65535) break;
    pl.NT = nt;
    pl.dual = (nt <= 64 && !a.dw_w) ? 1 : 0;  // small-N MMAs cost the same ~120 clk as N = 2*NT ones: fold two of the three
    pl.acc_cols = 32;
    while (pl.acc_cols < (pl.dual ? 2 * nt : nt)) pl.acc_cols <<= 1;
    pl.prm_floats = (a.dw_w ? 10 : 4) * a.CI + 3 * nt;
    pl.prm_floats = (pl.prm_floats + 3) & ~3;
    pl.dw_floats = a.dw_w ? ((2 * a.CI * 135 + 3) & ~3) : 0;
    const size_t misc = kSmemBars + (size_t)(pl.prm_floats + pl.dw_floats) * 4;
    const size_t w_all = (size_t)a.K * a.CI * nt * 4;
    const size_t x_all = (size_t)a.CI * pl.rows * 4;
    if (w_all + 2 * x_all + misc <= kSmemBudget) {  // weights resident, >= 2 input stages
      pl.resident = 1;
      pl.ci_chunk = a.CI;
      pl.n_chunks = 1;
      pl.wres_u4 = (int)(w_all / 16);
      pl.stage_u4 = (int)(x_all / 16);
      size_t ns = (kSmemBudget - misc - w_all) / x_all;
      pl.n_stages = (int)(ns > kMaxStages ? kMaxStages : ns);
      return true;
    }
    // stream weights with the input rows, per chunk of input channels
    for (int c = a.CI - a.CI % 16; c >= 16; c -= 16) {
      const size_t st = (size_t)c * ((size_t)pl.rows * 4 + (size_t)a.K * nt * 4);
      if (2 * st + misc <= kSmemBudget) {
        pl.resident = 0;
        pl.ci_chunk = c;
        pl.n_chunks = cdiv(a.CI, c);
        pl.wres_u4 = 0;
        pl.stage_u4 = (int)(st / 16);
        size_t ns = (kSmemBudget - misc) / st;
        pl.n_stages = (int)(ns > kMaxStages ? kMaxStages : ns);
        return true;
      }
    }
  }
  return false;
}

static int umma_in_mode(const sty_conv1d_args& a) {
  if (a.dw_w) {  // fused ConvNeXt front: pointwise conv only, nothing else in the prologue
    // the fused-front kernels are built LEAN (no residual / mask / scale / shuffle / sum in the epilogue)
    const bool ok = a.K == 1 && a.CI <= 64 && a.CI % 16 == 0 && a.dw_b && a.dw_gb &&
                    a.in_act == STY_ACT_NONE && !a.in_scale && !a.in_shift && !a.in_mask && !a.res &&
                    !a.out_mask && a.shuffle <= 1 && a.out_scale == 1.0f && !a.out_sum;
    return ok ? 4 : -1;
  }
  if (a.in_act == STY_ACT_SNAKE) return 3;
  if (a.in_act == STY_ACT_LEAKY02) return 2;
  if (a.in_act != STY_ACT_NONE) return 5;  // any other activation (GELU, ReLU, Swish): generic prologue, plain epilogue only
  return (a.in_scale || a.in_shift || a.in_mask) ? 1 : 0;
}
static int umma_out_mode(const sty_conv1d_args& a) {
  switch (a.out_act) {
    case STY_ACT_NONE: return 0;
    case STY_ACT_SNAKE: return 1;
    case STY_ACT_RELU: return 2;
    case STY_ACT_SWISH: return 3;
    default: return -1;
  }
}

bool conv1d_umma_eligible(const sty_conv1d_args& a) {
  if (!a.w_split || a.w_bs != 0 || a.T < 64) return false;
  if (umma_in_mode(a) < 0 || umma_out_mode(a) < 0) return false;
  if (umma_in_mode(a) == 5 && umma_out_mode(a) != 0) return false;
  if (a.out_sum && a.CO > 64) return false;
  UmmaPlan pl;
  return make_plan(a, pl);
}

// TMA staging of the raw input (see the TMA template parameter): needs 16-byte aligned rows (x_cs, x_bs
// multiples of 4 floats, aligned base), whole chunks, and room for >= 2 raw stages next to >= 2 operand stages.
static void plan_tma(const sty_conv1d_args& a, UmmaPlan& pl) {
  pl.tma = 0;
  pl.raw_rp = pl.raw_stages = pl.raw_u4 = pl.raw_shift = 0;
  static const bool off = getenv("STYLISH_B200_TMA") && atoi(getenv("STYLISH_B200_TMA")) == 0;
  if (off || a.dw_w || a.T < 512 || !tma_layout_ok(a.x, a.x_bs, a.x_cs)) return;
  const int shift = (4 - (a.pad & 3)) & 3;  // t0 is a multiple of 128: (t0 - pad - shift) % 4 == 0
  const int rp = (pl.rows + shift + 3) & ~3;
  if (rp > 256 || a.pad < 0) return;
  const size_t fixed = (size_t)pl.wres_u4 * 16 + (size_t)(pl.prm_floats + pl.dw_floats) * 4 + kSmemBars + 128;
  if (pl.resident) {
    // resident weights: the input may be staged in chunks of channels (the MMA loop offsets the weight rows) —
    // largest chunk that leaves room for 2 operand stages and 3 raw stages
    int c = a.CI;
    for (; c >= 16; c -= 16) {
      if (a.CI % c != 0 || c > 256) continue;
      if (fixed + 2 * ((size_t)c * pl.rows * 4) + 3 * ((size_t)c * rp * 4) <= kSmemBudget) break;
    }
    if (c < 16) return;
    pl.ci_chunk = c;
    pl.n_chunks = a.CI / c;
    pl.stage_u4 = (int)((size_t)c * pl.rows * 4 / 16);
  }
  if (a.CI % pl.ci_chunk != 0 || pl.ci_chunk > 256) return;
  const size_t raw_sz = (size_t)pl.ci_chunk * rp * 4;
  const size_t st = (size_t)pl.stage_u4 * 16;
  if (fixed + 2 * st + 2 * raw_sz > kSmemBudget) return;
  int ns = 2, nr = 2;
  while (nr < kMaxRaw && fixed + ns * st + (nr + 1) * raw_sz <= kSmemBudget) ++nr;
  while (ns < 3 && fixed + (ns + 1) * st + nr * raw_sz <= kSmemBudget) ++ns;
  pl.tma = 1;
  pl.raw_shift = shift;
  pl.raw_rp = rp;
  pl.raw_stages = nr;
  pl.raw_u4 = (int)(raw_sz / 16);
  pl.n_stages = ns;
}

using KernPtr = void (*)(const sty_conv1d_args, const UmmaPlan, const CUtensorMap);

template <bool TMA>
static KernPtr pick_kernel(const sty_conv1d_args& a) {
  const int im = umma_in_mode(a), om = umma_out_mode(a);
  // specialised epilogues of the S-rate generator convs (see EPI)
  const bool bare = !a.out_mask && a.shuffle <= 1 && a.out_scale == 1.0f;
  if (bare && !a.res) {
    if (im == 0 && om == 0) return conv1d_umma_kernel<0, 0, 1, TMA>;  // k21 input convs
    if (im == 0 && om == 1) return conv1d_umma_kernel<0, 1, 1, TMA>;  // pwconv1 + Snake (training graph)
    if (im == 3 && om == 0) return conv1d_umma_kernel<3, 0, 1, TMA>;  // AdaIN + Snake -> k11 (convs1)
  } else if (bare && a.res && a.res_scale == 1.0f) {
    if (im == 1 && om == 0) return conv1d_umma_kernel<1, 0, 2, TMA>;  // GRN scale -> pwconv2 + residual
    if (im == 3 && om == 0) return conv1d_umma_kernel<3, 0, 2, TMA>;  // AdaIN + Snake -> k11 + residual (convs2)
  }
  static const KernPtr table[4][4] = {
      {conv1d_umma_kernel<0, 0, 0, TMA>, conv1d_umma_kernel<0, 1, 0, TMA>, conv1d_umma_kernel<0, 2, 0, TMA>,
       conv1d_umma_kernel<0, 3, 0, TMA>},
      {conv1d_umma_kernel<1, 0, 0, TMA>, conv1d_umma_kernel<1, 1, 0, TMA>, conv1d_umma_kernel<1, 2, 0, TMA>,
       conv1d_umma_kernel<1, 3, 0, TMA>},
      {conv1d_umma_kernel<2, 0, 0, TMA>, conv1d_umma_kernel<2, 1, 0, TMA>, conv1d_umma_kernel<2, 2, 0, TMA>,
       conv1d_umma_kernel<2, 3, 0, TMA>},
      {conv1d_umma_kernel<3, 0, 0, TMA>, conv1d_umma_kernel<3, 1, 0, TMA>, conv1d_umma_kernel<3, 2, 0, TMA>,
       conv1d_umma_kernel<3, 3, 0, TMA>}};
  if (im >= 0 && im < 4) return table[im][om];
  if (im == 5 && om == 0) return conv1d_umma_kernel<5, 0, 0, TMA>;  // BatchNorm + GELU prologues (waveform discriminator)
  return nullptr;
}

int conv1d_umma_launch(const sty_conv1d_args& a, cudaStream_t st) {
  UmmaPlan pl;
  if (!make_plan(a, pl)) {
    set_error("conv1d_umma: shape not supported");
    return STY_ERR_BAD_ARG;
  }
  static int sms = 0;
  if (sms <= 0) {
    sms = sty_device_sm_count();
    if (sms <= 0) sms = 148;
  }
  plan_tma(a, pl);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (pl.tma && !make_tmap_bct(&tmap, a.x, a.B, a.CI, a.T, a.x_bs, a.x_cs, pl.raw_rp, pl.ci_chunk)) {
    UmmaPlan again;
    make_plan(a, again);
    pl = again;
    pl.tma = pl.raw_rp = pl.raw_stages = pl.raw_u4 = pl.raw_shift = 0;
  }
  const size_t smem = (size_t)pl.wres_u4 * 16 + (size_t)pl.n_stages * pl.stage_u4 * 16 +
                      (size_t)(pl.prm_floats + pl.dw_floats) * 4 + kSmemBars +
                      (pl.tma ? (size_t)pl.raw_stages * pl.raw_u4 * 16 + 128 : 128);
  KernPtr kern = nullptr;
  if (umma_in_mode(a) == 4) {
    static const KernPtr front[4] = {conv1d_umma_kernel<4, 0>, conv1d_umma_kernel<4, 1>, conv1d_umma_kernel<4, 2>,
                                     conv1d_umma_kernel<4, 3>};
    kern = front[umma_out_mode(a)];
  } else {
    kern = pl.tma ? pick_kernel<true>(a) : pick_kernel<false>(a);
  }
  if (!kern) {
    set_error("conv1d_umma: no kernel for this prologue / epilogue");
    return STY_ERR_BAD_ARG;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t n_tiles = (int64_t)a.B * pl.tiles_per_b;
  const int n_co = a.CO / pl.NT;
  int gx = (int)(n_tiles < sms ? n_tiles : sms);
  // several output-channel tiles: split the SMs between them (at least one CTA each)
  if (n_co > 1) {
    gx = sms / n_co;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = (int)n_tiles;
  }
  dim3 grid(gx, n_co, 1);
  kern<<<grid, pl.tma ? kThreadsTma : kThreads, smem, st>>>(a, pl, tmap);
  STY_CHECK_LAUNCH("conv1d_umma");
  return STY_OK;
}

}  // namespace sty
