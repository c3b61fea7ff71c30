// Conv1d weight gradient on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM).
//
//   dw[co, ci, k] = sum_{b,t} g[b,co,t] * u[b,ci,t + k*dil - pad]       (contraction over TIME)
//
// Both operands are activations staged channel-group-major, [8-channel group][time][8 ch] bf16, i.e. the
// UMMA canonical MN-major / no-swizzle layout with time as the K axis.  The M side of the MMA is
// built im2col-free: with a shared-memory descriptor whose 8-row-group stride (SBO) is `dil` time
// steps, the 16 row groups of one M=128 operand are the SAME staged 8 input channels seen through 16
// consecutive taps, so one MMA produces D[(tap, ci%8), co] for 16 taps at once and a k=21 conv needs
// 2 MMAs per 8 input channels and K step instead of 21.  (K == 1 has no taps: the M side is then the
// tensor with more channels, 128 per MMA.)  fp32 operands are split into bf16 (hi, lo) while staging and
// three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32, as in the forward kernel (conv1d_umma.cu).
//
// One CTA per SM walks a range of (batch, 128-step) chunks with a two-stage shared-memory ring and two
// roles handing over through mbarriers: 15 producer warps stage chunk i+1 (prologue / mask applied on the
// fly, 32 global loads in flight per thread) while a dedicated warp issues the MMAs of chunk i; the
// accumulators stay in TMEM for the whole range and are added to dw with atomics once.
#include "umma.cuh"

namespace sty {
namespace {

constexpr int kWgThreads = 480;  // 15 producer warps + 1 MMA warp = 4 warps per SM sub-partition: 128 registers each
constexpr int kTT = 128;  // time steps per chunk

struct WgPlan {
  int mode;        // 0: taps folded into the M side (K >= 8) | 1: K == 1, M side = input | 2: K == 1, M side = grad
                   // 3: 2 <= K < 8: M side = grad (128 channels), one accumulator per tap, the tap shift is the
                   //    start row of the N-side (input) descriptor
  int m_groups;    // staged 8-channel groups of the M side
  int n_groups;    // staged 8-channel groups of the N side (N = 8 * n_groups, multiple of 16)
  int rows_m;      // staged time steps of the M side
  int rows_n;      // staged time steps of the N side
  int tap_chunks;  // ceil(K / 16) in mode 0
  int n_acc;       // accumulators (N columns each)
  int acc_cols;    // TMEM columns allocated (power of two >= 32)
  int stages;      // shared-memory ring depth (1 or 2)
  int stage_u4;    // uint4 per stage
  int n_tchunks;   // ceil(T / 128)
  int m_tiles, n_tiles;
  int valid_rows;  // rows of the M side that carry data: 128 + (K-1)*dil
};

// One side of a chunk: `groups` 8-channel groups x `rows` time steps of the conv INPUT (prologue applied)
// or of the output GRADIENT (mask * scale applied), written as bf16 hi | lo planes
// dst[(split*groups + g)*rows + row].
struct SideDesc {
  uint4* dst;
  const float* src;   // batch element base
  const float* mask;  // (T) or nullptr
  int64_t cs;
  int C, c0, groups, rows, valid_rows, t_start, is_input;
  int blocked;  // 1: [time/8][group][8 steps] (core matrices of one K block adjacent), 0: [group][time]
};

// Stage one side of a chunk with `nthr` threads (this thread is `lt` of them).  Items = (8-channel group, time
// step) cells; a thread first issues the global loads of up to U items (32 loads in flight — staging is
// latency-bound otherwise), then transforms, splits and stores them.  (g, row) of successive items is stepped
// incrementally — no integer division in the loop.
// ACT : prologue activation known at compile time (STY_ACT_NONE, STY_ACT_SNAKE) or -1 = runtime switch.
// FAST: no mask and every staged channel exists, so only the per-item time-range predicate is left.  The generic body executes ~45 instructions per element
//       (ncu: 437 M warp instructions for 1.24 GB at 77 % i-cache hit rate, profiles/r01_ncu_full_wgrad_umma.txt);
//       the specialised bodies are what the S-rate weight gradients run.
template <bool IS_INPUT, int ACT, bool FAST>
__device__ __forceinline__ void stage_side_t(const SideDesc& S, const sty_conv1d_wgrad_args& p, const float* prm,
                                             int in_groups, int lt, int nthr) {
  constexpr int U = 4;  // 32 loads in flight per thread (6 measured slower)
  const int n_items = S.groups * S.rows;
  const int g_step = nthr / S.rows, r_step = nthr - g_step * S.rows;
  int g = lt / S.rows, row = lt - g * S.rows;
  const bool chan_full = FAST || S.c0 + S.groups * 8 <= S.C;  // every staged channel exists: no per-channel checks
  const int act = ACT >= 0 ? ACT : p.in_act;
  const int64_t cs = S.cs;
  for (int i0 = lt; i0 < n_items; i0 += nthr * U) {
    float v[U][8];
    int gg[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      gg[u] = g;
      rr[u] = row;
      const int t = S.t_start + row;
      const bool in_items = (i0 + u * nthr) < n_items;
      const bool ok = in_items && row < S.valid_rows && t >= 0 && t < p.T;
      const float* __restrict__ src = S.src + (int64_t)(S.c0 + g * 8) * cs + t;
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // running pointer: 2 integer instructions per load
        v[u][j] = (ok && (chan_full || S.c0 + g * 8 + j < S.C)) ? *src : 0.f;
        src += cs;
      }
      row += r_step;
      g += g_step;
      if (row >= S.rows) {
        row -= S.rows;
        ++g;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i0 + u * nthr >= n_items) continue;
      const int t = S.t_start + rr[u];
      const bool ok = rr[u] < S.valid_rows && t >= 0 && t < p.T;
      const float m = (!FAST && ok && S.mask) ? S.mask[t] : 1.f;
      if (IS_INPUT) {
        const int cl = gg[u] * 8;
        float sc[8], sh[8], al[8];
        *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(prm + cl);
        *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(prm + cl + 4);
        *reinterpret_cast<float4*>(sh) = *reinterpret_cast<const float4*>(prm + in_groups * 8 + cl);
        *reinterpret_cast<float4*>(sh + 4) = *reinterpret_cast<const float4*>(prm + in_groups * 8 + cl + 4);
        if (act == STY_ACT_SNAKE) {  // prm[2] = alpha, prm[3] = 1/alpha
          *reinterpret_cast<float4*>(al) = *reinterpret_cast<const float4*>(prm + 2 * in_groups * 8 + cl);
          *reinterpret_cast<float4*>(al + 4) = *reinterpret_cast<const float4*>(prm + 2 * in_groups * 8 + cl + 4);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float w = fmaf(FAST ? v[u][j] : v[u][j] * m, sc[j], sh[j]);
          if (act == STY_ACT_SNAKE) {
            const float sn = __sinf(al[j] * w);
            w += __fdividef(sn * sn, al[j]);
          } else if (act != STY_ACT_NONE) {
            w = act_apply(w, act);
          }
          v[u][j] = (ok && m >= 0.f && (chan_full || S.c0 + cl + j < S.C)) ? w : 0.f;  // m < 0: zero after the prologue
        }
      } else {
        const float mm = m * p.out_scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] *= mm;  // out-of-range cells were loaded as 0
      }
      uint32_t h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a0 = v[u][2 * j], a1 = v[u][2 * j + 1];
        h[j] = pack_bf16(a0, a1);
        l[j] = pack_bf16(a0 - __uint_as_float(h[j] << 16), a1 - __uint_as_float(h[j] & 0xffff0000u));
      }
      const int at = S.blocked ? (((rr[u] >> 3) * S.groups + gg[u]) << 3) + (rr[u] & 7) : gg[u] * S.rows + rr[u];
      S.dst[at] = make_uint4(h[0], h[1], h[2], h[3]);
      S.dst[S.groups * S.rows + at] = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

template <bool IS_INPUT>
__device__ __forceinline__ void stage_side(const SideDesc& S, const sty_conv1d_wgrad_args& p, const float* prm,
                                           int in_groups, int lt, int nthr) {
  const bool fast = !S.mask && S.c0 + S.groups * 8 <= S.C;
  if (!IS_INPUT) {
    if (fast) stage_side_t<false, STY_ACT_NONE, true>(S, p, prm, in_groups, lt, nthr);
    else stage_side_t<false, STY_ACT_NONE, false>(S, p, prm, in_groups, lt, nthr);
  } else if (p.in_act == STY_ACT_NONE && fast) {
    stage_side_t<true, STY_ACT_NONE, true>(S, p, prm, in_groups, lt, nthr);
  } else if (p.in_act == STY_ACT_SNAKE && fast) {
    stage_side_t<true, STY_ACT_SNAKE, true>(S, p, prm, in_groups, lt, nthr);
  } else if (p.in_act == STY_ACT_LEAKY02 && fast) {  // style-encoder ResBlk convs
    stage_side_t<true, STY_ACT_LEAKY02, true>(S, p, prm, in_groups, lt, nthr);
  } else {
    stage_side_t<true, -1, false>(S, p, prm, in_groups, lt, nthr);
  }
}

// Both sides of a chunk, concurrently: the producer threads are split between the two sides in proportion
// to their item counts (warp granularity), so a chunk costs one load round trip.
__device__ __forceinline__ void stage_chunk(const SideDesc& A, const SideDesc& Bd, const sty_conv1d_wgrad_args& p,
                                            const float* prm, int in_groups, int tid, int nthr_a) {
  if (tid < nthr_a) {
    if (A.is_input) stage_side<true>(A, p, prm, in_groups, tid, nthr_a);
    else stage_side<false>(A, p, prm, in_groups, tid, nthr_a);
  } else {
    if (Bd.is_input) stage_side<true>(Bd, p, prm, in_groups, tid - nthr_a, kWgThreads - nthr_a);
    else stage_side<false>(Bd, p, prm, in_groups, tid - nthr_a, kWgThreads - nthr_a);
  }
}

__global__ void __launch_bounds__(kWgThreads + 32, 1)
conv1d_wgrad_umma_kernel(const sty_conv1d_wgrad_args p, const WgPlan pl) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint4* stage0 = reinterpret_cast<uint4*>(smem_raw);
  float* prm = reinterpret_cast<float*>(stage0 + (size_t)pl.stages * pl.stage_u4);  // [3][8*input groups]
  const int in_groups = pl.mode >= 2 ? pl.n_groups : pl.m_groups;
  uint64_t* bars = reinterpret_cast<uint64_t*>(prm + ((3 * 8 * in_groups + 3) & ~3));
  uint64_t* full = bars;       // [2] producers -> MMA warp
  uint64_t* empty = bars + 2;  // [2] tensor core -> producers
  uint64_t* done = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = pl.n_groups * 8;

  const int m_tile = blockIdx.y / pl.n_tiles, n_tile = blockIdx.y % pl.n_tiles;
  const int m0 = m_tile * pl.m_groups * 8, n0 = n_tile * N;  // first channel of the M / N side
  const int ci0 = pl.mode >= 2 ? n0 : m0;                     // first INPUT channel of this tile

  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)pl.acc_cols);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], kWgThreads);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t total = (int64_t)p.B * pl.n_tchunks;
  const int64_t c_begin = (total * blockIdx.x) / gridDim.x, c_end = (total * (blockIdx.x + 1)) / gridDim.x;
  // instruction descriptor: D = f32, A = B = bf16, both MN-major, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int m_plane = pl.m_groups * pl.rows_m;  // uint4 per split plane of the M side
  const int n_plane = pl.n_groups * pl.rows_n;
  const uint32_t sbo_m = pl.mode == 0 ? (uint32_t)p.dil : (uint32_t)pl.rows_m;

  const uint32_t n_it = (uint32_t)(c_end - c_begin);
  // producer threads given to the M side: proportional to its share of the items, whole warps, >= 1 warp each
  int nthr_m;
  {
    const int im = pl.m_groups * pl.rows_m, in_ = pl.n_groups * pl.rows_n;
    int wm = (int)(((int64_t)(kWgThreads / 32) * im + (im + in_) / 2) / (im + in_));
    wm = wm < 1 ? 1 : (wm > kWgThreads / 32 - 1 ? kWgThreads / 32 - 1 : wm);
    nthr_m = wm * 32;
  }
  if (warp < kWgThreads / 32) {
    // =========================== producers (15 warps): stage chunk after chunk into the ring
    int cur_b = -1;
    uint32_t it = 0;
    for (int64_t chunk = c_begin; chunk < c_end; ++chunk, ++it) {
      const int b = (int)(chunk / pl.n_tchunks);
      const int t0 = (int)(chunk - (int64_t)b * pl.n_tchunks) * kTT;
      const uint32_t s = it % (uint32_t)pl.stages;
      mbar_wait(&empty[s], ((it / (uint32_t)pl.stages) & 1u) ^ 1u);  // MMAs that read this stage are done (no nanosleep: with a 2-stage ring an overslept wake-up stalls the tensor core)
      if (b != cur_b) {  // prologue parameters of this batch element
        asm volatile("bar.sync 1, %0;" ::"n"(kWgThreads) : "memory");
        for (int c = tid; c < 8 * in_groups; c += kWgThreads) {
          const int ci = ci0 + c;
          const bool ok = ci < p.CI;
          prm[c] = (ok && p.in_scale) ? p.in_scale[(int64_t)b * p.CI + ci] : 1.f;
          prm[8 * in_groups + c] = (ok && p.in_shift) ? p.in_shift[(int64_t)b * p.CI + ci] : 0.f;
          prm[16 * in_groups + c] = (ok && p.in_alpha) ? p.in_alpha[ci] : 1.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWgThreads) : "memory");
        cur_b = b;
      }
      uint4* Ms = stage0 + (size_t)s * pl.stage_u4;
      uint4* Ns = Ms + 2 * m_plane;
      const float* xb = p.x + (int64_t)b * p.x_bs;
      const float* gb = p.dy + (int64_t)b * p.dy_bs;
      const float* im = p.in_mask ? p.in_mask + (int64_t)b * p.T : nullptr;
      const float* om = p.out_mask ? p.out_mask + (int64_t)b * p.T : nullptr;
      SideDesc sm_, sn_;
      if (pl.mode >= 2) {  // M side = output gradient, N side = input
        sm_ = SideDesc{Ms, gb, om, p.dy_cs, p.CO, m0, pl.m_groups, pl.rows_m, kTT, t0, 0, pl.mode == 2};
        sn_ = SideDesc{Ns, xb, im, p.x_cs, p.CI, n0, pl.n_groups, pl.rows_n, pl.valid_rows, t0 - p.pad, 1,
                       pl.mode == 2};
      } else {             // M side = input (taps folded by the descriptor), N side = output gradient
        sm_ = SideDesc{Ms, xb, im, p.x_cs, p.CI, m0, pl.m_groups, pl.rows_m, pl.valid_rows, t0 - p.pad, 1,
                       pl.mode == 1};
        sn_ = SideDesc{Ns, gb, om, p.dy_cs, p.CO, n0, pl.n_groups, kTT, kTT, t0, 0, 1};
      }
      stage_chunk(sm_, sn_, p, prm, in_groups, tid, nthr_m);
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor-core proxy
      mbar_arrive(&full[s]);
    }
  } else {
    // =========================== MMA issuer (one elected lane of the extra warp)
    for (uint32_t it = 0; it < n_it; ++it) {
      const uint32_t s = it % (uint32_t)pl.stages;
      mbar_wait(&full[s], (it / (uint32_t)pl.stages) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint4* Ms = stage0 + (size_t)s * pl.stage_u4;
        const uint32_t m_addr = smem_u32(Ms), n_addr = smem_u32(Ms + 2 * m_plane);
        // blocked sides: K blocks are n_groups (m_groups) core matrices apart, groups adjacent (LBO = 8*groups, SBO = 8)
        const bool n_blk = pl.mode != 3, m_blk = pl.mode == 1 || pl.mode == 2;
        const uint64_t bd = n_blk ? make_desc(n_addr, 8u * (uint32_t)pl.n_groups, 8u)
                                  : make_desc(n_addr, 8u, (uint32_t)pl.rows_n);
        const uint32_t a_kstep = m_blk ? 16u * (uint32_t)pl.m_groups : 16u;  // 16 time steps, in 16-byte units
        const uint32_t b_kstep = n_blk ? 16u * (uint32_t)pl.n_groups : 16u;
        const uint32_t b_hi32 = (uint32_t)(bd >> 32);
        const uint32_t first = it == 0 ? 0u : 1u;
        for (int acc = 0; acc < pl.n_acc; ++acc) {
          const uint32_t b_lo = (uint32_t)bd + (pl.mode == 3 ? (uint32_t)(acc * p.dil) : 0u);  // mode 3: tap = acc
          // mode 0: accumulator (g, tc) = input-channel group g seen through taps [16 tc, 16 tc + 16)
          const int g = pl.mode == 0 ? acc / pl.tap_chunks : 0;
          const int tc = pl.mode == 0 ? acc - g * pl.tap_chunks : 0;
          const uint32_t a_off = (uint32_t)(g * pl.rows_m + tc * 16 * p.dil);  // uint4 units
          const uint64_t ad = m_blk ? make_desc(m_addr, 8u * (uint32_t)pl.m_groups, 8u)
                                    : make_desc(m_addr + a_off * 16u, 8u, sbo_m);
          const uint32_t a_hi32 = (uint32_t)(ad >> 32), a_lo = (uint32_t)ad;
          const uint32_t d = tmem_base + (uint32_t)(acc * N);
#pragma unroll 2
          for (uint32_t ks = 0; ks < kTT / 16; ++ks) {
            const uint32_t ak = a_lo + ks * a_kstep, bk = b_lo + ks * b_kstep;
            umma_bf16_w(d, ak, a_hi32, bk, b_hi32, idesc, ks == 0 ? first : 1u);            // hi * hi
            umma_bf16_w(d, ak + (uint32_t)m_plane, a_hi32, bk, b_hi32, idesc, 1u);          // lo * hi
            umma_bf16_w(d, ak, a_hi32, bk + (uint32_t)n_plane, b_hi32, idesc, 1u);          // hi * lo
          }
        }
        umma_commit(&empty[s]);                      // stage free once the tensor core has read it
        if (it == n_it - 1) umma_commit(done);      // ... and every accumulator is complete
      }
      __syncwarp();
    }
  }
  // everyone waits for the last MMA before the accumulators are read back
  const uint32_t it = n_it;
  if (n_it > 0) mbar_wait_sleep(done, 0);
  tc_fence_after();

  // ---- accumulators -> dw (atomics).  thread owns TMEM lane m = 32*(warp&3) + lane; the four warp
  // groups split the 16-column chunks
  constexpr int kParts = (kWgThreads / 32) / 4;  // complete groups of 4 warps (one per TMEM lane quadrant)
  if (it > 0 && warp < 4 * kParts) {
    const int q = warp & 3, part = warp >> 2;
    const int m = q * 32 + lane;
    const int chunks16 = (pl.n_acc * N) >> 4;
    for (int c = part; c < chunks16; c += kParts) {
      float r[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r);
      const int col = c * 16;
      const int acc = col / N, nn = col - acc * N;
      int ci, co_base, tap;
      bool row_ok;
      if (pl.mode == 0) {
        const int g = acc / pl.tap_chunks, tc = acc - g * pl.tap_chunks;
        tap = tc * 16 + (m >> 3);
        ci = m0 + g * 8 + (m & 7);
        co_base = n0 + nn;
        row_ok = tap < p.K && ci < p.CI;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (row_ok && co_base + j < p.CO)
            atomicAdd(p.dw + ((int64_t)(co_base + j) * p.CI + ci) * p.K + tap, r[j]);
      } else if (pl.mode == 3) {
        const int co = m0 + m;
        const int ci_base = n0 + nn;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co < p.CO && ci_base + j < p.CI)
            atomicAdd(p.dw + ((int64_t)co * p.CI + ci_base + j) * p.K + acc, r[j]);
      } else if (pl.mode == 1) {
        ci = m0 + m;
        co_base = n0 + nn;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (ci < p.CI && co_base + j < p.CO) atomicAdd(p.dw + (int64_t)(co_base + j) * p.CI + ci, r[j]);
      } else {
        const int co = m0 + m;
        const int ci_base = n0 + nn;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (co < p.CO && ci_base + j < p.CI) atomicAdd(p.dw + (int64_t)co * p.CI + ci_base + j, r[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)pl.acc_cols);
}

constexpr size_t kWgSmemBudget = 220 * 1024;

bool make_wg_plan(const sty_conv1d_wgrad_args& a, WgPlan& pl) {
  if (a.CI % 8 != 0 || a.CO % 8 != 0 || a.K > 32 || a.dil > 64) return false;
  if (a.in_act != STY_ACT_NONE && a.in_act != STY_ACT_SNAKE && a.in_act != STY_ACT_LEAKY02 &&
      a.in_act != STY_ACT_RELU && a.in_act != STY_ACT_SWISH && a.in_act != STY_ACT_GELU)
    return false;
  pl.n_tchunks = cdiv(a.T, kTT);
  if (a.K == 1) {
    pl.mode = a.CO >= a.CI ? 2 : 1;
    const int Cm = pl.mode == 2 ? a.CO : a.CI, Cn = pl.mode == 2 ? a.CI : a.CO;
    pl.m_groups = 16;
    pl.rows_m = pl.rows_n = kTT;
    pl.valid_rows = kTT;
    pl.tap_chunks = 1;
    pl.n_acc = 1;
    int ng = cdiv(Cn, 8);
    if (ng & 1) ++ng;
    if (ng > 32) ng = 32;
    pl.n_groups = ng;
    pl.m_tiles = cdiv(Cm, 128);
    pl.n_tiles = cdiv(Cn, 8 * ng);
  } else if (a.K < 8) {
    pl.mode = 3;
    pl.m_groups = 16;
    pl.rows_m = kTT;
    pl.valid_rows = pl.rows_n = kTT + (a.K - 1) * a.dil;
    pl.tap_chunks = 1;
    pl.n_acc = a.K;
    int ng = (512 / a.K) / 16 * 2;  // N = 8*ng: multiple of 16, K accumulators within 512 TMEM columns
    if (ng > 32) ng = 32;
    int need = cdiv(a.CI, 8);
    if (need & 1) ++need;
    if (ng > need) ng = need;
    pl.n_groups = ng;
    pl.m_tiles = cdiv(a.CO, 128);
    pl.n_tiles = cdiv(a.CI, 8 * ng);
  } else {
    pl.mode = 0;
    pl.rows_n = kTT;
    pl.tap_chunks = cdiv(a.K, 16);
    pl.valid_rows = kTT + (a.K - 1) * a.dil;
    pl.rows_m = kTT + (16 * pl.tap_chunks - 1) * a.dil;
    int ng = cdiv(a.CO, 8);
    if (ng & 1) ++ng;
    if (ng > 32) ng = 32;
    pl.n_groups = ng;
    const int acc_max = 512 / (8 * ng);
    int mg = acc_max / pl.tap_chunks;
    if (mg < 1) return false;
    if (mg > a.CI / 8) mg = a.CI / 8;
    pl.m_groups = mg;
    pl.n_acc = mg * pl.tap_chunks;
    pl.m_tiles = cdiv(a.CI / 8, mg);
    pl.n_tiles = cdiv(a.CO, 8 * ng);
  }
  pl.acc_cols = 32;
  while (pl.acc_cols < pl.n_acc * pl.n_groups * 8) pl.acc_cols <<= 1;
  if (pl.acc_cols > 512) return false;
  pl.stage_u4 = 2 * (pl.m_groups * pl.rows_m + pl.n_groups * pl.rows_n);
  const int in_groups = pl.mode >= 2 ? pl.n_groups : pl.m_groups;
  const size_t misc = (size_t)((3 * 8 * in_groups + 3) & ~3) * 4 + 128;
  const size_t st = (size_t)pl.stage_u4 * 16;
  if (2 * st + misc <= kWgSmemBudget) pl.stages = 2;
  else if (st + misc <= kWgSmemBudget) pl.stages = 1;
  else return false;
  if ((int64_t)pl.m_tiles * pl.n_tiles > 65535) return false;
  return true;
}

}  // namespace

bool conv1d_wgrad_umma_eligible(const sty_conv1d_wgrad_args& a) {
  WgPlan pl;
  return a.tensor_cores != 0 && make_wg_plan(a, pl);
}

int conv1d_wgrad_umma_launch(const sty_conv1d_wgrad_args& a, cudaStream_t st) {
  WgPlan pl;
  if (!make_wg_plan(a, pl)) {
    set_error("conv1d_wgrad_umma: shape not supported");
    return STY_ERR_BAD_ARG;
  }
  static int sms = 0;
  if (sms <= 0) {
    sms = sty_device_sm_count();
    if (sms <= 0) sms = 148;
  }
  const int in_groups = pl.mode >= 2 ? pl.n_groups : pl.m_groups;
  const size_t smem = (size_t)pl.stages * pl.stage_u4 * 16 + (size_t)((3 * 8 * in_groups + 3) & ~3) * 4 + 128;
  cudaFuncSetAttribute(conv1d_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = pl.m_tiles * pl.n_tiles;
  const int64_t total = (int64_t)a.B * pl.n_tchunks;
  int64_t gx = sms / tiles;
  if (gx < 1) gx = 1;
  if (gx > total) gx = total;
  dim3 grid((unsigned)gx, (unsigned)tiles);
  conv1d_wgrad_umma_kernel<<<grid, kWgThreads + 32, smem, st>>>(a, pl);
  STY_CHECK_LAUNCH("conv1d_wgrad_umma");
  return STY_OK;
}

}  // namespace sty
