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
#include <cuda_bf16.h>

#include "common.cuh"

namespace sty {

// ----------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]),
        "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]),
        "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, sm_100 version bit.
// lbo / sbo in units of 16 bytes (K-direction / 8-row-group strides).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo16, uint32_t sbo16) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) |
         ((uint64_t)(sbo16 & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// -------------------------------------------------------------------- kernel
// grid: (ceil(T/128), CO/NT, B); 128 threads.  Thread t of the CTA owns accumulator row t;
// NT (<= 256, multiple of 16) output channels per CTA.
template <bool PRO>
__global__ void __launch_bounds__(128)
conv1d_umma_kernel(const sty_conv1d_args p, const int ci_chunk, const int rows, const int NT,
                   const uint32_t tmem_cols) {
  constexpr int MT = 128;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int K = p.K, dil = p.dil, CO = p.CO, CI = p.CI;
  const int co0 = blockIdx.y * NT;
  const int c8n = ci_chunk >> 3;                       // 16-byte K chunks per staged chunk
  uint4* Xs = reinterpret_cast<uint4*>(smem_raw);      // [2][c8n][rows]
  uint4* Ws = Xs + 2 * c8n * rows;                     // [K][2][c8n][NT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(Ws + (size_t)K * 2 * c8n * NT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * MT;

  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const float* __restrict__ xb = p.x + (int64_t)b * p.x_bs;
  const uint4* __restrict__ wsplit = reinterpret_cast<const uint4*>(p.w_split);
  const float* __restrict__ in_mask = (PRO && p.in_mask) ? p.in_mask + (int64_t)b * p.T : nullptr;
  const float* __restrict__ in_scale = (PRO && p.in_scale) ? p.in_scale + (int64_t)b * CI : nullptr;
  const float* __restrict__ in_shift = (PRO && p.in_shift) ? p.in_shift + (int64_t)b * CI : nullptr;
  const float* __restrict__ in_alpha = (PRO && p.in_alpha) ? p.in_alpha : nullptr;
  const int in_act = PRO ? p.in_act : STY_ACT_NONE;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N = CO, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);

  uint32_t phase = 0, accumulate = 0;
  for (int c0 = 0; c0 < CI; c0 += ci_chunk) {
    const int cc8 = min(ci_chunk, CI - c0) >> 3;
    // ---- weights of this chunk: K*2 blocks of cc8*CO 16-byte vectors (pre-split bf16 hi/lo)
    {
      const int blk_elems = cc8 * NT;
      const int total = K * 2 * blk_elems;
      for (int idx = tid; idx < total; idx += MT) {
        const int blk = idx / blk_elems, within = idx - blk * blk_elems;
        const int c8l = within / NT, n = within - c8l * NT;
        Ws[blk * (c8n * NT) + within] =
            wsplit[((int64_t)blk * (CI >> 3) + (c0 >> 3) + c8l) * CO + co0 + n];
      }
    }
    // ---- input rows: 8 channels x 1 time step per item -> two 16-byte vectors (hi, lo)
    for (int c8 = 0; c8 < cc8; ++c8) {
      const int cbase = c0 + c8 * 8;
      float sc[8], sh[8], al[8], ia[8];
      if constexpr (PRO) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sc[j] = in_scale ? in_scale[cbase + j] : 1.f;
          sh[j] = in_shift ? in_shift[cbase + j] : 0.f;
          al[j] = in_alpha ? in_alpha[cbase + j] : 1.f;
          ia[j] = 1.0f / al[j];
        }
      }
      for (int row = tid; row < rows; row += MT) {
        const int t = t0 - p.pad + row;
        const bool ok = (t >= 0) && (t < p.T);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ok ? xb[(int64_t)(cbase + j) * p.x_cs + t] : 0.f;
        if constexpr (PRO) {
          const float m = (ok && in_mask) ? in_mask[t] : 1.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float w = fmaf(v[j] * m, sc[j], sh[j]);
            if (in_act == STY_ACT_SNAKE) {
              w = fmaf(ia[j], sin_sq(al[j] * w), w);
            } else if (in_act == STY_ACT_LEAKY02) {
              w = w > 0.f ? w : 0.2f * w;
            } else if (in_act != STY_ACT_NONE) {
              w = act_apply(w, in_act);
            }
            v[j] = ok ? w : 0.f;
          }
        }
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hi[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
          lo[j] = v[j] - hi[j];
        }
        uint4 h4, l4;
        h4.x = pack_bf16(hi[0], hi[1]); h4.y = pack_bf16(hi[2], hi[3]);
        h4.z = pack_bf16(hi[4], hi[5]); h4.w = pack_bf16(hi[6], hi[7]);
        l4.x = pack_bf16(lo[0], lo[1]); l4.y = pack_bf16(lo[2], lo[3]);
        l4.z = pack_bf16(lo[4], lo[5]); l4.w = pack_bf16(lo[6], lo[7]);
        Xs[(0 * c8n + c8) * rows + row] = h4;
        Xs[(1 * c8n + c8) * rows + row] = l4;
      }
    }
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t xs_addr = smem_u32(Xs), ws_addr = smem_u32(Ws);
      const int kblocks = cc8 >> 1;  // MMA K = 16 bf16 = two 16-byte chunks
      // (x split, w split): hi*hi, lo*hi, hi*lo
#pragma unroll 1
      for (int combo = 0; combo < 3; ++combo) {
        const int sx = (combo == 1) ? 1 : 0, sw = (combo == 2) ? 1 : 0;
#pragma unroll 1
        for (int tap = 0; tap < K; ++tap) {
#pragma unroll 1
          for (int kb = 0; kb < kblocks; ++kb) {
            const uint32_t a_addr = xs_addr + (uint32_t)(((sx * c8n + 2 * kb) * rows + tap * dil) * 16);
            const uint32_t b_addr = ws_addr + (uint32_t)((((tap * 2 + sw) * c8n + 2 * kb) * NT) * 16);
            umma_bf16(tmem_base, make_desc(a_addr, (uint32_t)rows, 8u), make_desc(b_addr, (uint32_t)NT, 8u),
                      idesc, accumulate);
            accumulate = 1;
          }
        }
      }
      umma_commit(bar);  // arrives on `bar` when every MMA above has finished reading smem
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    accumulate = 1;
  }

  // ---- epilogue: thread tid owns output time step t0 + tid (TMEM lane tid)
  const int t = t0 + tid;
  const bool t_ok = t < p.T;
  const float* __restrict__ out_mask = p.out_mask ? p.out_mask + (int64_t)b * p.T : nullptr;
  float* __restrict__ yb = p.y + (int64_t)b * p.y_bs;
  const float* __restrict__ rb = p.res ? p.res + (int64_t)b * p.r_bs : nullptr;
  const int s = p.shuffle > 1 ? p.shuffle : 1;
  const int out_act = p.out_act;
  const float om = ((out_mask && t_ok) ? out_mask[t] : 1.f) * p.out_scale;
  for (int n0 = 0; n0 < NT; n0 += 32) {
    float r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, r);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int co = co0 + n0 + j;
      if (n0 + j < NT) {  // NT % 16 == 0: the tail chunk may be half used
        const float bias = p.bias ? p.bias[co] : 0.f;
        float v = r[j] + bias;
        if (out_act == STY_ACT_SNAKE) {
          const float al = p.out_alpha[co];
          v = fmaf(1.0f / al, sin_sq(al * v), v);
        } else if (out_act != STY_ACT_NONE) {
          v = act_apply(v, out_act);
        }
        v *= om;
        float sq = 0.f;
        if (t_ok) {
          const int c_out = co / s, r_out = co - c_out * s;
          const int64_t off = (int64_t)t * s + r_out;
          if (rb) v = fmaf(p.res_scale, rb[(int64_t)c_out * p.r_cs + off], v);
          yb[(int64_t)c_out * p.y_cs + off] = v;
          sq = v * v;
        }
        if (p.out_sumsq) {
          sq = warp_sum(sq);
          if (lane == 0) atomicAdd(p.out_sumsq + (int64_t)b * CO + co, sq);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// smem bytes for a chunk of `chunk` input channels
static size_t umma_smem_bytes(int chunk, int rows, int K, int NT) {
  return (size_t)chunk * rows * 4 + (size_t)K * chunk * NT * 4 + 16;
}

// output-channel tile: all of CO when <= 256, else the largest multiple-of-16 divisor <= 256
static int umma_co_tile(int CO) {
  if (CO <= 256) return CO;
  for (int n = 256; n >= 16; n -= 16)
    if (CO % n == 0) return n;
  return 0;
}

bool conv1d_umma_eligible(const sty_conv1d_args& a) {
  if (!a.w_split) return false;
  if (a.CI % 16 != 0 || a.CO % 16 != 0 || a.CO < 16) return false;
  if (a.w_bs != 0 || a.T < 128) return false;
  const int nt = umma_co_tile(a.CO);
  if (nt < 16 || a.CO / nt > 65535) return false;
  const int rows = 128 + (a.K - 1) * a.dil;
  return umma_smem_bytes(16, rows, a.K, nt) <= 200 * 1024 && rows < 16384;
}

int conv1d_umma_launch(const sty_conv1d_args& a, cudaStream_t st) {
  const int rows = 128 + (a.K - 1) * a.dil;
  const int nt = umma_co_tile(a.CO);
  // largest chunk (multiple of 16) whose footprint allows two CTAs per SM, else one
  int chunk = 16;
  for (int c = a.CI - a.CI % 16; c >= 16; c -= 16) {
    if (umma_smem_bytes(c, rows, a.K, nt) <= 100 * 1024) {
      chunk = c;
      break;
    }
  }
  const size_t smem = umma_smem_bytes(chunk, rows, a.K, nt);
  uint32_t cols = 32;
  while ((int)cols < nt) cols <<= 1;
  const bool pro = a.in_scale || a.in_shift || a.in_mask || a.in_act != STY_ACT_NONE;
  auto kern = pro ? conv1d_umma_kernel<true> : conv1d_umma_kernel<false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(a.T, 128), a.CO / nt, a.B);
  kern<<<grid, 128, smem, st>>>(a, chunk, rows, nt, cols);
  STY_CHECK_LAUNCH("conv1d_umma");
  return STY_OK;
}

}  // namespace sty
