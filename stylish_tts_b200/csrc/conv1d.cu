// Generic stride-1 Conv1d, im2col-free, fp32 FMA, with fused prologue/epilogue.
//
// One CTA computes a CO_TILE x T_TILE output tile of one batch element.  The
// input rows (CI chunk x (T_TILE + halo)) are staged ONCE in shared memory with
// the prologue (mask, AdaIN/GRN affine, LeakyReLU/Snake) applied while staging,
// so normalisation + activation cost no extra HBM pass; the weight chunk is
// staged next to it in (ci, k, co) order so each thread reads its CO_R weights
// as broadcast 128-bit loads.  Threads own a CO_R x T_R register tile whose T
// positions are interleaved by the lane count (lane-consecutive shared loads,
// fully coalesced global stores).
//
// Semantics: see sty_conv1d_fwd in include/stylish_b200.h.
#include "common.cuh"

namespace sty {

template <int CO_TILE, int T_TILE, int CO_R, int T_R, int KT>
__global__ void __launch_bounds__((CO_TILE / CO_R) * (T_TILE / T_R))
conv1d_kernel(const sty_conv1d_args p, const int ci_chunk, const int xtp) {
  constexpr int TL = T_TILE / T_R;  // threads along time
  constexpr int NT = (CO_TILE / CO_R) * TL;
  static_assert(TL <= 32 && (TL & (TL - 1)) == 0, "TL must be a power of two <= 32");
  static_assert(CO_R % 4 == 0, "CO_R must be a multiple of 4");
  const int K = KT > 0 ? KT : p.K;
  const int dil = p.dil;
  const int XT = T_TILE + (K - 1) * dil;

  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                   // [ci_chunk][xtp]
  float* ws = smem + ci_chunk * xtp;  // [ci_chunk][K][CO_TILE]

  const int b = blockIdx.z;
  const int co0 = blockIdx.y * CO_TILE;
  const int t0 = blockIdx.x * T_TILE;
  const int tid = threadIdx.x;
  const int tl = tid % TL;
  const int cg = tid / TL;

  float acc[CO_R][T_R];
#pragma unroll
  for (int i = 0; i < CO_R; ++i)
#pragma unroll
    for (int j = 0; j < T_R; ++j) acc[i][j] = 0.f;

  const float* __restrict__ xb = p.x + (int64_t)b * p.x_bs;
  const float* __restrict__ wb = p.w + (int64_t)b * p.w_bs;
  const float* __restrict__ in_mask = p.in_mask ? p.in_mask + (int64_t)b * p.T : nullptr;
  const bool co_vec = (p.CO % 4 == 0);

  for (int ci0 = 0; ci0 < p.CI; ci0 += ci_chunk) {
    const int cc = min(ci_chunk, p.CI - ci0);
    // ---- stage the input rows, prologue applied, zero outside [0,T)
    for (int ci = 0; ci < cc; ++ci) {
      const int c = ci0 + ci;
      const float* __restrict__ xr = xb + (int64_t)c * p.x_cs;
      float sc = 1.f, sh = 0.f, al = 1.f;
      if (p.in_scale) sc = p.in_scale[(int64_t)b * p.CI + c];
      if (p.in_shift) sh = p.in_shift[(int64_t)b * p.CI + c];
      if (p.in_alpha) al = p.in_alpha[c];
      for (int tt = tid; tt < xtp; tt += NT) {
        const int t = t0 - p.pad + tt;
        float v = 0.f;
        if (tt < XT && t >= 0 && t < p.T) {
          v = xr[t];
          if (in_mask) v *= in_mask[t];
          v = fmaf(v, sc, sh);
          v = act_apply(v, p.in_act, al);
        }
        xs[ci * xtp + tt] = v;
      }
    }
    // ---- stage the weight chunk (rows of CO_TILE contiguous floats)
    {
      const int rows = cc * K;
      const float* __restrict__ wsrc = wb + (int64_t)ci0 * K * p.CO + co0;
      if (co_vec) {
        constexpr int V = CO_TILE / 4;
        for (int idx = tid; idx < rows * V; idx += NT) {
          const int r = idx / V, c4 = (idx - r * V) * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (co0 + c4 < p.CO) v = *reinterpret_cast<const float4*>(wsrc + (int64_t)r * p.CO + c4);
          *reinterpret_cast<float4*>(ws + r * CO_TILE + c4) = v;
        }
      } else {
        for (int idx = tid; idx < rows * CO_TILE; idx += NT) {
          const int r = idx / CO_TILE, c = idx - r * CO_TILE;
          ws[idx] = (co0 + c < p.CO) ? wsrc[(int64_t)r * p.CO + c] : 0.f;
        }
      }
    }
    __syncthreads();
    // ---- accumulate
    for (int ci = 0; ci < cc; ++ci) {
      const float* __restrict__ xrow = xs + ci * xtp + tl;
      const float* __restrict__ wrow = ws + ci * K * CO_TILE + cg * CO_R;
      auto mac = [&](const int k) {
        float wv[CO_R], xv[T_R];
#pragma unroll
        for (int i = 0; i < CO_R; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wrow + k * CO_TILE + i);
          wv[i] = w4.x; wv[i + 1] = w4.y; wv[i + 2] = w4.z; wv[i + 3] = w4.w;
        }
#pragma unroll
        for (int j = 0; j < T_R; ++j) xv[j] = xrow[k * dil + j * TL];
#pragma unroll
        for (int i = 0; i < CO_R; ++i)
#pragma unroll
          for (int j = 0; j < T_R; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
      };
      if constexpr (KT > 0) {
#pragma unroll
        for (int k = 0; k < KT; ++k) mac(k);
      } else {
        for (int k = 0; k < K; ++k) mac(k);
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  const float* __restrict__ out_mask = p.out_mask ? p.out_mask + (int64_t)b * p.T : nullptr;
  float* __restrict__ yb = p.y + (int64_t)b * p.y_bs;
  const float* __restrict__ rb = p.res ? p.res + (int64_t)b * p.r_bs : nullptr;
  const int s = p.shuffle > 1 ? p.shuffle : 1;
#pragma unroll
  for (int i = 0; i < CO_R; ++i) {
    const int co = co0 + cg * CO_R + i;
    const bool co_ok = co < p.CO;
    const float bias = (co_ok && p.bias) ? p.bias[co] : 0.f;
    const float al = (co_ok && p.out_alpha) ? p.out_alpha[co] : 1.f;
    float ssq = 0.f;
    if (co_ok) {
      const int c_out = co / s, r_out = co - c_out * s;
#pragma unroll
      for (int j = 0; j < T_R; ++j) {
        const int t = t0 + tl + j * TL;
        if (t < p.T) {
          float v = acc[i][j] + bias;
          v = act_apply(v, p.out_act, al);
          if (out_mask) v *= out_mask[t];
          v *= p.out_scale;
          const int64_t off = (int64_t)c_out * p.y_cs + (int64_t)t * s + r_out;
          if (rb) v = fmaf(p.res_scale, rb[(int64_t)c_out * p.r_cs + (int64_t)t * s + r_out], v);
          yb[off] = v;
          ssq = fmaf(v, v, ssq);
        }
      }
    }
    if (p.out_sumsq) {
#pragma unroll
      for (int o = TL / 2; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
      if (tl == 0 && co_ok) atomicAdd(p.out_sumsq + (int64_t)b * p.CO + co, ssq);
    }
  }
}

template <int CO_TILE, int T_TILE, int CO_R, int T_R, int KT>
static int launch_cfg(const sty_conv1d_args& a, cudaStream_t st) {
  constexpr int NT = (CO_TILE / CO_R) * (T_TILE / T_R);
  const int XT = T_TILE + (a.K - 1) * a.dil;
  const int xtp = (XT + 3) & ~3;
  const int per_ci = (xtp + a.K * CO_TILE) * (int)sizeof(float);
  int chunk = (40 * 1024) / per_ci;
  if (chunk < 1) chunk = 1;
  if (chunk > 32) chunk = 32;
  if (chunk > a.CI) chunk = a.CI;
  const size_t smem = (size_t)chunk * per_ci;
  auto kern = conv1d_kernel<CO_TILE, T_TILE, CO_R, T_R, KT>;
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) {
      set_error("conv1d: kernel footprint too large (K=%d dil=%d)", a.K, a.dil);
      return STY_ERR_BAD_ARG;
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  dim3 grid(cdiv(a.T, T_TILE), cdiv(a.CO, CO_TILE), a.B);
  kern<<<grid, NT, smem, st>>>(a, chunk, xtp);
  STY_CHECK_LAUNCH("conv1d");
  return STY_OK;
}

template <int CO_TILE, int T_TILE, int CO_R, int T_R>
static int launch_k(const sty_conv1d_args& a, cudaStream_t st) {
  switch (a.K) {
    case 1: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 1>(a, st);
    case 3: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 3>(a, st);
    case 5: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 5>(a, st);
    case 11: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 11>(a, st);
    case 21: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 21>(a, st);
    default: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 0>(a, st);
  }
}

}  // namespace sty

extern "C" int sty_conv1d_fwd(const sty_conv1d_args* a, sty_stream_t stream) {
  using namespace sty;
  STY_REQUIRE(a != nullptr, "conv1d: null args");
  STY_REQUIRE(a->x && a->w && a->y, "conv1d: null tensor pointer");
  STY_REQUIRE(a->B > 0 && a->CI > 0 && a->CO > 0 && a->T > 0, "conv1d: bad shape B=%d CI=%d CO=%d T=%d",
              a->B, a->CI, a->CO, a->T);
  STY_REQUIRE(a->K >= 1 && a->K <= 64 && a->dil >= 1 && a->pad >= 0, "conv1d: bad K=%d dil=%d pad=%d",
              a->K, a->dil, a->pad);
  STY_REQUIRE(2 * a->pad == (a->K - 1) * a->dil, "conv1d: only 'same' padding is supported (K=%d dil=%d pad=%d)",
              a->K, a->dil, a->pad);
  STY_REQUIRE(a->in_act != STY_ACT_SNAKE || a->in_alpha, "conv1d: snake prologue needs in_alpha");
  STY_REQUIRE(a->out_act != STY_ACT_SNAKE || a->out_alpha, "conv1d: snake epilogue needs out_alpha");
  STY_REQUIRE(a->shuffle <= 1 || a->CO % a->shuffle == 0, "conv1d: CO %% shuffle != 0");
  STY_REQUIRE(a->shuffle <= 1 || a->out_sumsq == nullptr, "conv1d: sumsq with shuffle unsupported");
  cudaStream_t st = as_stream(stream);
  if (a->CO <= 32) {
    if (a->T > 96) return launch_k<32, 256, 8, 8>(*a, st);
    return launch_k<32, 64, 8, 4>(*a, st);
  }
  if (a->T > 96) return launch_k<64, 128, 8, 8>(*a, st);
  return launch_k<64, 64, 8, 4>(*a, st);
}
