// Generic stride-1 Conv1d, im2col-free, fp32 FMA, with fused prologue/epilogue.
//
// One CTA computes a CO_TILE x T_TILE output tile of one batch element.  The
// input rows (CI chunk x (T_TILE + halo)) are staged ONCE in shared memory with
// the prologue (mask, AdaIN/GRN affine, LeakyReLU/Snake) applied while staging,
// so normalisation + activation cost no extra HBM pass; the weight chunk is
// staged next to it in (ci, k, co) order so each thread reads its CO_R weights
// as broadcast 128-bit loads.  Threads own a CO_R x T_R register tile whose T
// positions are interleaved by the lane count (lane-consecutive shared loads,
// fully coalesced global stores).  Staging loads are issued 8 rows at a time so
// each thread keeps 8+ global loads in flight.
//
// Semantics: see sty_conv1d_fwd in include/stylish_b200.h.
#include "common.cuh"

namespace sty {

template <int CO_TILE, int T_TILE, int CO_R, int T_R, int KT, bool PRO>
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
  const float* __restrict__ in_mask = (PRO && p.in_mask) ? p.in_mask + (int64_t)b * p.T : nullptr;
  const float* __restrict__ in_scale = (PRO && p.in_scale) ? p.in_scale + (int64_t)b * p.CI : nullptr;
  const float* __restrict__ in_shift = (PRO && p.in_shift) ? p.in_shift + (int64_t)b * p.CI : nullptr;
  const float* __restrict__ in_alpha = (PRO && p.in_alpha) ? p.in_alpha : nullptr;
  const int in_act = PRO ? p.in_act : STY_ACT_NONE;
  const bool co_vec = (p.CO % 4 == 0);

  for (int ci0 = 0; ci0 < p.CI; ci0 += ci_chunk) {
    const int cc = min(ci_chunk, p.CI - ci0);
    // ---- stage the input rows, prologue applied, zero outside [0,T)
    for (int tt = tid; tt < xtp; tt += NT) {
      const int t = t0 - p.pad + tt;
      const bool ok = (tt < XT) && (t >= 0) && (t < p.T);
      const float* __restrict__ xr = xb + (int64_t)ci0 * p.x_cs + t;
      float* __restrict__ xd = xs + tt;
      float m = 1.f;
      if (PRO && ok && in_mask) m = in_mask[t];
      constexpr int U = 8;  // rows per batch: all loads of a batch are issued before any use
      for (int cb = 0; cb < cc; cb += U) {
        float v[U], sc[U], sh[U], al[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int ci = cb + u;
          v[u] = (ok && ci < cc) ? xr[(int64_t)ci * p.x_cs] : 0.f;
        }
        if constexpr (PRO) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int c = min(ci0 + cb + u, p.CI - 1);
            sc[u] = in_scale ? in_scale[c] : 1.f;
            sh[u] = in_shift ? in_shift[c] : 0.f;
            al[u] = in_alpha ? in_alpha[c] : 1.f;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float w = fmaf(v[u] * m, sc[u], sh[u]);
            if (in_act == STY_ACT_SNAKE) {
              w = fmaf(1.0f / al[u], sin_sq(al[u] * w), w);
            } else if (in_act == STY_ACT_LEAKY02) {
              w = w > 0.f ? w : 0.2f * w;
            } else if (in_act != STY_ACT_NONE) {
              w = act_apply(w, in_act);
            }
            v[u] = (ok && m >= 0.f) ? w : 0.f;  // negative mask value: zero AFTER the prologue
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (cb + u < cc) xd[(cb + u) * xtp] = v[u];
      }
    }
    // ---- stage the weight chunk (rows of CO_TILE contiguous floats)
    {
      const int rows = cc * K;
      const float* __restrict__ wsrc = wb + (int64_t)ci0 * K * p.CO + co0;
      if (co_vec) {
        constexpr int V = CO_TILE / 4;
#pragma unroll 4
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
  const int out_act = p.out_act;
  float om[T_R];
#pragma unroll
  for (int j = 0; j < T_R; ++j) {
    const int t = t0 + tl + j * TL;
    om[j] = (out_mask && t < p.T) ? out_mask[t] : 1.f;
    om[j] *= p.out_scale;
  }
#pragma unroll
  for (int i = 0; i < CO_R; ++i) {
    const int co = co0 + cg * CO_R + i;
    const bool co_ok = co < p.CO;
    const float bias = (co_ok && p.bias) ? p.bias[co] : 0.f;
    const float al = (co_ok && p.out_alpha) ? p.out_alpha[co] : 1.f;
    const float inv_al = 1.0f / al;
    float ssq = 0.f, ssum = 0.f;
    if (co_ok) {
      const int c_out = co / s, r_out = co - c_out * s;
      float* __restrict__ yrow = yb + (int64_t)c_out * p.y_cs + r_out;
      const float* __restrict__ rrow = rb ? rb + (int64_t)c_out * p.r_cs + r_out : nullptr;
#pragma unroll
      for (int j = 0; j < T_R; ++j) {
        const int t = t0 + tl + j * TL;
        if (t < p.T) {
          float v = acc[i][j] + bias;
          v = act_apply(v, out_act, al, inv_al);
          v *= om[j];
          if (rrow) v = fmaf(p.res_scale, rrow[(int64_t)t * s], v);
          yrow[(int64_t)t * s] = v;
          ssq = fmaf(v, v, ssq);
          ssum += v;
        }
      }
    }
    if (p.out_sumsq) {
#pragma unroll
      for (int o = TL / 2; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
      if (tl == 0 && co_ok) atomicAdd(p.out_sumsq + (int64_t)b * p.CO + co, ssq);
    }
    if (p.out_sum) {
#pragma unroll
      for (int o = TL / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      if (tl == 0 && co_ok) atomicAdd(p.out_sum + (int64_t)b * p.CO + co, ssum);
    }
  }
}

bool conv1d_umma_eligible(const sty_conv1d_args& a);         // conv1d_umma.cu
int conv1d_umma_launch(const sty_conv1d_args& a, cudaStream_t st);

template <int CO_TILE, int T_TILE, int CO_R, int T_R, int KT, bool PRO>
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
  auto kern = conv1d_kernel<CO_TILE, T_TILE, CO_R, T_R, KT, PRO>;
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

template <int CO_TILE, int T_TILE, int CO_R, int T_R, bool PRO>
static int launch_k(const sty_conv1d_args& a, cudaStream_t st) {
  switch (a.K) {
    case 1: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 1, PRO>(a, st);
    case 3: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 3, PRO>(a, st);
    case 5: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 5, PRO>(a, st);
    case 11: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 11, PRO>(a, st);
    case 21: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 21, PRO>(a, st);
    default: return launch_cfg<CO_TILE, T_TILE, CO_R, T_R, 0, PRO>(a, st);
  }
}

template <int CO_TILE, int T_TILE, int CO_R, int T_R>
static int launch_p(const sty_conv1d_args& a, cudaStream_t st) {
  const bool pro = a.in_scale || a.in_shift || a.in_mask || a.in_act != STY_ACT_NONE;
  return pro ? launch_k<CO_TILE, T_TILE, CO_R, T_R, true>(a, st)
             : launch_k<CO_TILE, T_TILE, CO_R, T_R, false>(a, st);
}

// Tile choice: minimise (#waves x tile work), i.e. prefer the largest tile that still
// fills the 148 SMs for about two waves; small T / few channels fall to small tiles.
static int pick_config(const sty_conv1d_args& a) {
  static int sms = 0;
  if (sms <= 0) {
    sms = sty_device_sm_count();
    if (sms <= 0) sms = 148;
  }
  //                co_tile t_tile threads  resident CTAs per SM (registers/threads)
  const int cfg[5][4] = {{128, 128, 256, 2}, {64, 128, 128, 4}, {32, 256, 128, 4},
                         {64, 64, 128, 4},   {32, 64, 64, 8}};
  double best = 1e30;
  int pick = 4;
  for (int i = 0; i < 5; ++i) {
    const int cot = cfg[i][0], tt = cfg[i][1], thr = cfg[i][2], occ = cfg[i][3];
    if (cot > 32 && cot > ((a.CO + 31) / 32) * 32) continue;  // tile wider than the output
    const int64_t blocks = (int64_t)cdiv(a.T, tt) * cdiv(a.CO, cot) * a.B;
    const int64_t per_sm = (blocks + sms - 1) / sms;  // CTAs the busiest SM executes
    const double resident = (double)(per_sm < occ ? per_sm : occ) * thr / 32.0;
    const double eff = resident >= 8.0 ? 1.0 : resident / 8.0;  // too few warps: latency-bound
    double cost = (double)per_sm * cot * tt / eff;
    cost *= 1.0 + 4.0 / cot + 8.0 / tt;  // smaller tiles re-read more weights / inputs
    if (cost < best) {
      best = cost;
      pick = i;
    }
  }
  return pick;
}

}  // namespace sty

extern "C" int sty_conv1d_fwd(const sty_conv1d_args* a, sty_stream_t stream) {
  using namespace sty;
  STY_REQUIRE(a != nullptr, "conv1d: null args");
  STY_REQUIRE(a->x && a->w && a->y, "conv1d: null tensor pointer");
  STY_REQUIRE(a->B > 0 && a->CI > 0 && a->CO > 0 && a->T > 0, "conv1d: bad shape B=%d CI=%d CO=%d T=%d",
              a->B, a->CI, a->CO, a->T);
  STY_REQUIRE(a->B <= 65535, "conv1d: batch too large for the grid");
  STY_REQUIRE(a->K >= 1 && a->K <= 64 && a->dil >= 1 && a->pad >= 0, "conv1d: bad K=%d dil=%d pad=%d",
              a->K, a->dil, a->pad);
  STY_REQUIRE(2 * a->pad == (a->K - 1) * a->dil, "conv1d: only 'same' padding is supported (K=%d dil=%d pad=%d)",
              a->K, a->dil, a->pad);
  STY_REQUIRE(a->in_act != STY_ACT_SNAKE || a->in_alpha, "conv1d: snake prologue needs in_alpha");
  STY_REQUIRE(a->out_act != STY_ACT_SNAKE || a->out_alpha, "conv1d: snake epilogue needs out_alpha");
  STY_REQUIRE(a->shuffle <= 1 || a->CO % a->shuffle == 0, "conv1d: CO %% shuffle != 0");
  STY_REQUIRE(a->shuffle <= 1 || (a->out_sumsq == nullptr && a->out_sum == nullptr),
              "conv1d: sum / sumsq with shuffle unsupported");
  cudaStream_t st = as_stream(stream);
  if (conv1d_umma_eligible(*a)) return conv1d_umma_launch(*a, st);
  STY_REQUIRE(a->dw_w == nullptr, "conv1d: the fused ConvNeXt front needs the tensor-core path "
                                  "(K=1, CI<=64, CI,CO %% 16 == 0, T>=128, w_split set)");
  switch (pick_config(*a)) {
    case 0: return launch_p<128, 128, 8, 8>(*a, st);
    case 1: return launch_p<64, 128, 8, 8>(*a, st);
    case 2: return launch_p<32, 256, 8, 8>(*a, st);
    case 3: return launch_p<64, 64, 8, 4>(*a, st);
    default: return launch_p<32, 64, 8, 4>(*a, st);
  }
}
