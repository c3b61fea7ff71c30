// Dense token GEMMs of the style-diffusion denoiser (BASELINE configs[3]) on the 5th-gen tensor
// cores, fed by TMA:   C[m, n] = act( sum_k A[m, k] W[n, k] + bias[n] ) (+ residual[m, n])
//
// Precision: "bf16x3" like the convolutions — every fp32 operand is kept in memory as TWO bf16 planes (hi, lo) and
// three MMAs (hi*hi + lo*hi + hi*lo) accumulate in fp32 in TMEM.  Because the planes are produced by the epilogue
// of the PREVIOUS kernel (GEMM, LayerNorm, attention), the operand tiles go global -> shared by cp.async.bulk.tensor
// straight into the UMMA K-major SWIZZLE_128B layout, and the MMA warp consumes them with no thread ever touching
// the data.
//
// CTA: 128 x 256 output tile (N = 256 makes the MMA tensor-bound: 128 cycles vs 96 of shared-memory operand reads),
// K blocks of 64 (= one 128-byte swizzle atom per row), 2-stage ring of 96 KB (A hi, A lo: 128 rows; W hi, W lo: 256
// rows), tiles in the canonical SWIZZLE_128B K-major layout written by TMA with full 128-byte rows, persistent.
// warp 0: TMA loader | warp 1: MMA issuer (elect.sync, 12 MMAs per K block) | warps 2-5: epilogue (TMEM -> bias /
// GELU / residual -> fp32 rows and / or bf16 hi|lo planes), 2 accumulator stages of 256 TMEM columns.
#include <string.h>

#include "tma.cuh"

namespace sty {
namespace {

constexpr int kBM = 128, kBN = 256, kBK = 64;
constexpr int kStages = 2;
constexpr int kAU4 = kBM * (kBK / 8);  // uint4 per A plane tile: 128 rows x 128 bytes
constexpr int kWU4 = kBN * (kBK / 8);  // uint4 per W plane tile: 256 rows x 128 bytes
constexpr int kStageU4 = 2 * kAU4 + 2 * kWU4;
constexpr int kStageBytes = kStageU4 * 16;
constexpr int kAccCols = kBN;

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 bytes apart, sm_100 version bit
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
constexpr int kGemmThreads = 6 * 32;

struct GemmArgs {
  const float* bias;      // (N) or null
  const float* res;       // (M, N) fp32 or null
  float* out;             // (M, N) fp32 or null
  __nv_bfloat16* out_split;  // [2][M][N] or null
  int M, N, K;            // M % 128 == 0 (rows beyond the real M are padding), N % 128 == 0, K % 64 == 0
  int act;                // STY_ACT_NONE | STY_ACT_GELU
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16x3_kernel(const GemmArgs p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint4* stage0 = reinterpret_cast<uint4*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 1) tmem_alloc(tmem_slot, 2 * kAccCols);
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128);
    }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_tiles_n = p.N / kBN, n_tiles = (p.M / kBM) * n_tiles_n, kblocks = p.K / kBK;

  if (warp == 0) {
    // =========================== TMA loader
    if (elect_one()) {
      prefetch_tensormap(&tmA);
      prefetch_tensormap(&tmW);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * kBM, n0 = (tile % n_tiles_n) * kBN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
          mbar_wait_sleep(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], kStageBytes);
          uint4* st = stage0 + (size_t)s * kStageU4;
          // planes are stacked along the row axis of the tensor map: rows [0, M) = hi, [M, 2M) = lo
          tma_load_2d(st, &tmA, &full[s], kb * kBK, m0);
          tma_load_2d(st + kAU4, &tmA, &full[s], kb * kBK, p.M + m0);
          tma_load_2d(st + 2 * kAU4, &tmW, &full[s], kb * kBK, n0);
          tma_load_2d(st + 2 * kAU4 + kWU4, &tmW, &full[s], kb * kBK, p.N + n0);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    uint32_t it = 0, j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const uint32_t a = j & 1u;
      mbar_wait(&acc_empty[a], ((j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + a * (uint32_t)kAccCols;
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const uint32_t s = it % kStages, ph = (it / kStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint4* st = stage0 + (size_t)s * kStageU4;
          const uint64_t ad = make_desc_sw128(smem_u32(st));
          const uint64_t wd = make_desc_sw128(smem_u32(st + 2 * kAU4));
          const uint32_t ah = (uint32_t)(ad >> 32), wh = (uint32_t)(wd >> 32);
          const uint32_t al = (uint32_t)ad, wl = (uint32_t)wd;
#pragma unroll
          for (uint32_t ks = 0; ks < kBK / 16; ++ks) {  // a 16-element K step = 32 bytes inside the swizzle atom
            const uint32_t ak = al + ks * 2u, wk = wl + ks * 2u;
            umma_bf16_w(d, ak, ah, wk, wh, idesc, (kb | (int)ks) ? 1u : 0u);   // A hi * W hi
            umma_bf16_w(d, ak + kAU4, ah, wk, wh, idesc, 1u);                   // A lo * W hi
            umma_bf16_w(d, ak, ah, wk + kWU4, wh, idesc, 1u);                   // A hi * W lo
          }
          umma_commit(&empty[s]);
          if (kb == kblocks - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue: thread = output row (TMEM lane), 4 chunks of 32 columns
    const int q = warp & 3;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int m0 = (tile / n_tiles_n) * kBM, n0 = (tile % n_tiles_n) * kBN;
      const uint32_t a = j & 1u;
      const int m = m0 + q * 32 + lane;
      mbar_wait_sleep(&acc_full[a], (j >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < kBN / 32; ++ch) {
        float r[32];
        tmem_ld32(tmem_base + a * (uint32_t)kAccCols + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
        if (ch == kBN / 32 - 1) {
          tc_fence_before();
          mbar_arrive(&acc_empty[a]);
        }
        const int n = n0 + ch * 32;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + n + i);
            r[i] += b.x, r[i + 1] += b.y, r[i + 2] += b.z, r[i + 3] += b.w;
          }
        }
        if (p.act == STY_ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0.5f * r[i] * (1.0f + erff(r[i] * 0.70710678118654752440f));
        }
        if (p.res) {
          const float* rr = p.res + (int64_t)m * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(rr + i);
            r[i] += b.x, r[i + 1] += b.y, r[i + 2] += b.z, r[i + 3] += b.w;
          }
        }
        if (p.out) {
          float* o = p.out + (int64_t)m * p.N + n;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(r[i], r[i + 1], r[i + 2], r[i + 3]);
        }
        if (p.out_split) {
          uint4* oh = reinterpret_cast<uint4*>(p.out_split + (int64_t)m * p.N + n);
          uint4* ol = reinterpret_cast<uint4*>(p.out_split + ((int64_t)p.M + m) * p.N + n);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v0 = r[g * 8 + 2 * e], v1 = r[g * 8 + 2 * e + 1];
              h[e] = pack_bf16(v0, v1);
              l[e] = pack_bf16(v0 - __uint_as_float(h[e] << 16), v1 - __uint_as_float(h[e] & 0xffff0000u));
            }
            oh[g] = make_uint4(h[0], h[1], h[2], h[3]);
            ol[g] = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * kAccCols);
}

// ---- small token-rate helpers --------------------------------------------------------------------------------
// fp32 (n) -> bf16 hi | lo planes (out[0..n) = hi, out[n..2n) = lo)
__global__ void split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (int64_t)gridDim.x * blockDim.x * 2) {
    const float a = x[i], b = (i + 1 < n) ? x[i + 1] : 0.f;
    const uint32_t h = pack_bf16(a, b);
    const uint32_t l = pack_bf16(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xffff0000u));
    if (i + 1 < n) {
      *reinterpret_cast<uint32_t*>(out + i) = h;
      *reinterpret_cast<uint32_t*>(out + n + i) = l;
    } else {
      out[i] = __ushort_as_bfloat16((unsigned short)(h & 0xffff));
      out[n + i] = __ushort_as_bfloat16((unsigned short)(l & 0xffff));
    }
  }
}

// tokens[m, :] = [ scale * x[b, 0:Cx] | emb[b, t, 0:Ce] ],  m = b * T + t   (rows >= B*T: zeros)
__global__ void build_tokens_kernel(const float* __restrict__ x, const float* __restrict__ emb, float scale,
                                    float* __restrict__ tok, int B, int T, int Cx, int Ce, int M) {
  const int C = Cx + Ce;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)M * C; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / C), c = (int)(i - (int64_t)m * C);
    float v = 0.f;
    if (m < B * T) {
      const int b = m / T;
      v = c < Cx ? scale * x[(int64_t)b * Cx + c] : emb[(int64_t)m * Ce + (c - Cx)];
    }
    tok[i] = v;
  }
}

// hm[m, :] = h[m, :] + add[b, :];  n = LN(hm) * gamma + beta  -> bf16 hi | lo planes.  One warp per row, C % 128 == 0.
template <int C>
__global__ void __launch_bounds__(256)
row_ln_split_kernel(const float* __restrict__ h, const float* __restrict__ add, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, float* __restrict__ hm, __nv_bfloat16* __restrict__ out,
                    int M, int M_real, int T) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int PER = C / 128;  // float4 per lane
  float4 v[PER];
  const bool real = row < M_real;
  const float* ar = add ? add + (int64_t)(real ? row / T : 0) * C : nullptr;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = (i * 32 + lane) * 4;
    v[i] = real ? *reinterpret_cast<const float4*>(h + (int64_t)row * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (ar && real) {
      const float4 a = *reinterpret_cast<const float4*>(ar + c);
      v[i].x += a.x, v[i].y += a.y, v[i].z += a.z, v[i].w += a.w;
    }
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s * (1.0f / C);
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
    qv += a * a + b * b + c2 * c2 + d * d;
  }
  for (int off = 16; off; off >>= 1) qv += __shfl_xor_sync(0xffffffffu, qv, off);
  const float rstd = rsqrtf(qv * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (hm) *reinterpret_cast<float4*>(hm + (int64_t)row * C + c) = v[i];
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
    float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
    if (!real) y0 = y1 = y2 = y3 = 0.f;
    const uint32_t h0 = pack_bf16(y0, y1), h1 = pack_bf16(y2, y3);
    const uint32_t l0 = pack_bf16(y0 - __uint_as_float(h0 << 16), y1 - __uint_as_float(h0 & 0xffff0000u));
    const uint32_t l1 = pack_bf16(y2 - __uint_as_float(h1 << 16), y3 - __uint_as_float(h1 & 0xffff0000u));
    *reinterpret_cast<uint2*>(out + (int64_t)row * C + c) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(out + ((int64_t)M + row) * C + c) = make_uint2(l0, l1);
  }
}

// out[b, c] = mean_t x[b*T + t, c]
__global__ void token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int C) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += x[((int64_t)b * T + t) * C + c];
  out[(int64_t)b * C + c] = s / (float)T;
}

bool make_tmap_planes(CUtensorMap* out, const void* base, int64_t rows, int64_t K, int box_rows) {
  // row-major bf16 [rows, K]: box (64 K-elements = 128 bytes, box_rows), SWIZZLE_128B
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return false;
    fn = reinterpret_cast<Fn>(ptr);
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace sty

using namespace sty;

extern "C" int sty_split_planes_fwd(const float* x, void* out, int64_t n, sty_stream_t stream) {
  STY_REQUIRE(x && out && n > 0 && (n & 1) == 0, "split_planes: bad argument (n must be even)");
  int64_t blocks = (n / 2 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_planes_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(out), n);
  STY_CHECK_LAUNCH("split_planes");
  return STY_OK;
}

extern "C" int sty_gemm_split_fwd(const void* a_split, const void* w_split, const float* bias, const float* res,
                                  float* out, void* out_split, int M, int N, int K, int act, sty_stream_t stream) {
  STY_REQUIRE(a_split && w_split && (out || out_split), "gemm_split: null pointer");
  STY_REQUIRE(M > 0 && M % kBM == 0 && N > 0 && N % kBN == 0 && K > 0 && K % kBK == 0,
              "gemm_split: M %% 128, N %% 256, K %% 64 == 0 required (got %d %d %d)", M, N, K);
  STY_REQUIRE(act == STY_ACT_NONE || act == STY_ACT_GELU, "gemm_split: activation not supported");
  CUtensorMap tmA, tmW;
  STY_REQUIRE(make_tmap_planes(&tmA, a_split, 2 * (int64_t)M, K, kBM) &&
                  make_tmap_planes(&tmW, w_split, 2 * (int64_t)N, K, kBN),
              "gemm_split: tensor map encoding failed");
  GemmArgs g;
  g.bias = bias; g.res = res; g.out = out; g.out_split = reinterpret_cast<__nv_bfloat16*>(out_split);
  g.M = M; g.N = N; g.K = K; g.act = act;
  static int sms = 0;
  if (sms <= 0) {
    sms = sty_device_sm_count();
    if (sms <= 0) sms = 148;
  }
  const int n_tiles = (M / kBM) * (N / kBN);
  const size_t smem = (size_t)kStages * kStageBytes + 256;
  cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  gemm_bf16x3_kernel<<<n_tiles < sms ? n_tiles : sms, kGemmThreads, smem, as_stream(stream)>>>(g, tmA, tmW);
  STY_CHECK_LAUNCH("gemm_split");
  return STY_OK;
}

extern "C" int sty_build_tokens_fwd(const float* x, const float* emb, float scale, float* tok, int B, int T, int Cx,
                                    int Ce, int M, sty_stream_t stream) {
  STY_REQUIRE(x && emb && tok && B > 0 && T > 0 && M >= B * T, "build_tokens: bad argument");
  build_tokens_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(x, emb, scale, tok, B, T, Cx, Ce, M);
  STY_CHECK_LAUNCH("build_tokens");
  return STY_OK;
}

extern "C" int sty_row_ln_split_fwd(const float* h, const float* add, const float* gamma, const float* beta, float eps,
                                    float* hm, void* out_split, int M, int M_real, int T, int C,
                                    sty_stream_t stream) {
  STY_REQUIRE(h && gamma && beta && out_split && M > 0 && M_real <= M && T > 0, "row_ln_split: bad argument");
  STY_REQUIRE(C == 1024, "row_ln_split: built for 1024 features (got %d)", C);
  row_ln_split_kernel<1024><<<(M + 7) / 8, 256, 0, as_stream(stream)>>>(
      h, add, gamma, beta, eps, hm, reinterpret_cast<__nv_bfloat16*>(out_split), M, M_real, T);
  STY_CHECK_LAUNCH("row_ln_split");
  return STY_OK;
}

extern "C" int sty_token_mean_fwd(const float* x, float* out, int B, int T, int C, sty_stream_t stream) {
  STY_REQUIRE(x && out && B > 0 && T > 0 && C > 0, "token_mean: bad argument");
  dim3 grid((C + 127) / 128, B);
  token_mean_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, out, T, C);
  STY_CHECK_LAUNCH("token_mean");
  return STY_OK;
}
