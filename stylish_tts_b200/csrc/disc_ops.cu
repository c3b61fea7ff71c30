// Spectrogram-discriminator kernels that do not belong on the tensor cores (reference
// train/models/discriminator.py:13-69).  Images are "row-channel": (B, Hp = bins + 2, C, W) with zero border rows,
// element (b, r, c, w) at ((b*Hp + r)*C + c)*W + w.
//
//   first layer   Conv2d(1 -> 32, 3x9, pad (1,4)) on the raw spectrogram (B, bins, W): 864 FMAs per pixel, output-write
//                 bound; forward, data gradient (generator half only) and weight gradient.
//   layer "tail"  everything that hangs off a 32-channel pre-activation h besides the next 32 -> 32 convolution:
//                 a = LeakyReLU_0.1(h); score = Conv2d(32 -> 1, 3x3)(a); next = a or space-to-depth(a) for the following
//                 stride-(1,2) layer.  One pass reads h and writes score + next; the backward reads h, d(next), d(score)
//                 and writes dh = leaky'(h) * (score-conv data gradient + d(next)) — the activation `a` itself is never
//                 stored and the two gradient contributions are never added by a separate pass.
//
// All fp32 FMA (the 32 -> 32 3x9 / 3x3 convolutions in between run on tcgen05: conv1d_umma.cu through RowConvFn).
#include "common.cuh"

namespace sty {
namespace {

constexpr int kC = 32;        // channels of every hidden layer
constexpr float kSlope = 0.1f;

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : kSlope * v; }

// ------------------------------------------------------------------------------------------------ first layer, forward
// block: one image row (b, r), 512 columns; thread t owns columns w0 + t and w0 + t + 256 for all 32 channels
constexpr int kF_TW = 512, kF_Threads = 256;
__global__ void __launch_bounds__(kF_Threads)
disc_first_fwd_kernel(const float* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ h, int bins, int W) {
  __shared__ __align__(16) float ws[kC * 28];
  __shared__ float ys[3][kF_TW + 8];
  const int Hp = bins + 2;
  const int b = blockIdx.z, r = blockIdx.y, w0 = blockIdx.x * kF_TW, tid = threadIdx.x;
  float* __restrict__ hrow = h + ((int64_t)(b * Hp + r) * kC) * W;
  if (r == 0 || r == Hp - 1) {  // zero border rows
    for (int c = 0; c < kC; ++c)
      for (int j = tid; j < kF_TW; j += kF_Threads)
        if (w0 + j < W) hrow[(int64_t)c * W + w0 + j] = 0.f;
    return;
  }
  for (int i = tid; i < kC * 28; i += kF_Threads) {
    const int c = i / 28, k = i % 28;
    ws[i] = k < 27 ? w[c * 27 + k] : bias[c];
  }
  for (int i = tid; i < 3 * (kF_TW + 8); i += kF_Threads) {
    const int dr = i / (kF_TW + 8), j = i % (kF_TW + 8);
    const int rho = r - 1 + dr - 1, ww = w0 + j - 4;  // spectrogram row of image row r + dr - 1
    ys[dr][j] = (rho >= 0 && rho < bins && ww >= 0 && ww < W) ? y[((int64_t)b * bins + rho) * W + ww] : 0.f;
  }
  __syncthreads();
  float x0[27], x1[27];
#pragma unroll
  for (int dr = 0; dr < 3; ++dr)
#pragma unroll
    for (int dk = 0; dk < 9; ++dk) {
      x0[dr * 9 + dk] = ys[dr][tid + dk];
      x1[dr * 9 + dk] = ys[dr][tid + 256 + dk];
    }
  const int wa = w0 + tid, wb = w0 + tid + 256;
#pragma unroll 2
  for (int c = 0; c < kC; ++c) {
    const float4* wv = reinterpret_cast<const float4*>(ws + c * 28);
    float a0, a1;
    float wk[28];
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float4 v = wv[q];
      wk[4 * q] = v.x; wk[4 * q + 1] = v.y; wk[4 * q + 2] = v.z; wk[4 * q + 3] = v.w;
    }
    a0 = a1 = wk[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      a0 = fmaf(wk[k], x0[k], a0);
      a1 = fmaf(wk[k], x1[k], a1);
    }
    if (wa < W) hrow[(int64_t)c * W + wa] = a0;
    if (wb < W) hrow[(int64_t)c * W + wb] = a1;
  }
}

// ---------------------------------------------------------------------------------------- first layer, weight gradient
// dw[c,dr,dk] += sum y[b, rho + dr - 1, w + dk - 4] * dh[b, rho + 1, c, w];  db[c] += sum dh.
// persistent blocks; lane = channel, warp = 32-column slice of a 256-column tile; accumulators live in registers
// across all tiles of the block, one flush (shared-memory reduction over the warps + 896 atomics) per block.
constexpr int kW_TW = 256, kW_Threads = 256, kW_Pitch = kW_TW + 1;
__global__ void __launch_bounds__(kW_Threads)
disc_first_wgrad_kernel(const float* __restrict__ y, const float* __restrict__ dh, float* __restrict__ dw,
                        float* __restrict__ db, int B, int bins, int W) {
  extern __shared__ __align__(16) float sm[];
  float* dhs = sm;                         // [32][kW_Pitch]
  float* ys = sm + kC * kW_Pitch;          // [3][kW_TW + 8]  (16-byte aligned: 32*257*4 % 16 == 0)
  const int Hp = bins + 2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wt = (W + kW_TW - 1) / kW_TW;
  const int64_t tiles = (int64_t)B * bins * wt;
  float acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.f;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int xt = (int)(tile % wt);
    const int rho = (int)((tile / wt) % bins), b = (int)(tile / ((int64_t)wt * bins));
    const int w0 = xt * kW_TW;
    __syncthreads();
    const float* __restrict__ src = dh + ((int64_t)(b * Hp + rho + 1) * kC) * W;
    for (int i = tid; i < kC * kW_TW; i += kW_Threads) {
      const int c = i / kW_TW, j = i % kW_TW;
      dhs[c * kW_Pitch + j] = (w0 + j < W) ? src[(int64_t)c * W + w0 + j] : 0.f;
    }
    for (int i = tid; i < 3 * (kW_TW + 8); i += kW_Threads) {
      const int dr = i / (kW_TW + 8), j = i % (kW_TW + 8);
      const int rr = rho + dr - 1, ww = w0 + j - 4;
      ys[i] = (rr >= 0 && rr < bins && ww >= 0 && ww < W) ? y[((int64_t)b * bins + rr) * W + ww] : 0.f;
    }
    __syncthreads();
    // this warp: columns warp*32 .. +31, 8 at a time (y window of 16 values per row in registers)
#pragma unroll 1
    for (int j0 = warp * 32; j0 < warp * 32 + 32; j0 += 8) {
      float g[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        g[q] = dhs[lane * kW_Pitch + j0 + q];
        acc[27] += g[q];
      }
#pragma unroll
      for (int dr = 0; dr < 3; ++dr) {
        const float4* yv = reinterpret_cast<const float4*>(ys + dr * (kW_TW + 8) + j0);
        const float4 a = yv[0], bq = yv[1], cq = yv[2], dq = yv[3];
        const float v[16] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, cq.x, cq.y, cq.z, cq.w, dq.x, dq.y, dq.z, dq.w};
#pragma unroll
        for (int dk = 0; dk < 9; ++dk)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[dr * 9 + dk] = fmaf(v[q + dk], g[q], acc[dr * 9 + dk]);
      }
    }
  }
  // flush: sum the 8 warps through shared memory
  __syncthreads();
  float* red = sm;  // [8][32][28]
#pragma unroll
  for (int k = 0; k < 28; ++k) red[(warp * 32 + lane) * 28 + k] = acc[k];
  __syncthreads();
  for (int i = tid; i < kC * 28; i += kW_Threads) {
    float s = 0.f;
    for (int q = 0; q < 8; ++q) s += red[q * kC * 28 + i];
    const int c = i / 28, k = i % 28;
    if (k < 27) atomicAdd(dw + c * 27 + k, s);
    else atomicAdd(db + c, s);
  }
}

// ----------------------------------------------------------------------------- row-streaming 32 -> 1 convolution (+ tail)
// One kernel serves the two places where 32 channels are contracted into one map by a 3 x KW 'same' convolution:
//   tail forward      KW = 3, LEAKY: a = LeakyReLU(h); score = conv(a) + bias; next = a | space-to-depth(a) | nothing
//   first-layer dgrad KW = 9, FLIP : dy = conv(dh) with the taps of the 1 -> 32 forward kernel reversed on both axes
// block: (b, strip of kT_RS rows, 128 columns).  Rows stream through a 3-slot shared-memory ring; the NEXT row's global
// loads are issued into registers before the current row is consumed (latency hidden behind the FMAs), `next` is
// written straight from those registers.  Per row: warp q contracts channels 4q..4q+3 for 4 adjacent columns per lane,
// the 8 partial maps are summed through shared memory.  2 barriers per row.
constexpr int kT_TW = 128, kT_Threads = 256, kT_RS = 16;
template <int NEXT, int KW, bool LEAKY, bool FLIP>
__global__ void __launch_bounds__(kT_Threads)
disc_rows_to_map_kernel(const float* __restrict__ h, const float* __restrict__ wsc, const float* __restrict__ bsc,
                        float* __restrict__ score, float* __restrict__ next, int Hp, int W) {
  constexpr int HALO = KW / 2, COLS = kT_TW + 2 * HALO, PITCH = (COLS + 3) & ~3;  // tile column j holds w0 + j - HALO
  constexpr int NQ = (COLS + 31) / 32, WK = (3 * KW + 3) & ~3;
  extern __shared__ __align__(16) float sm[];
  float* ring = sm;                                 // [3][32][PITCH]
  float* wsm = sm + 3 * kC * PITCH;                 // [32][WK]
  float* part = wsm + kC * WK;                      // [8][128]
  const int b = blockIdx.z, r0 = blockIdx.y * kT_RS, w0 = blockIdx.x * kT_TW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bins = Hp - 2, W2 = (W + 1) >> 1;
  for (int i = tid; i < kC * WK; i += kT_Threads) {
    const int cc = i / WK, k = i % WK;
    float v = 0.f;
    if (k < 3 * KW) v = FLIP ? wsc[cc * 3 * KW + (2 - k / KW) * KW + (KW - 1 - k % KW)] : wsc[cc * 3 * KW + k];
    wsm[i] = v;
  }
  const float bias = bsc ? bsc[0] : 0.f;
  const int r_end = min(r0 + kT_RS, Hp);
  float pre[4][NQ];
  auto fetch = [&](int rr) {
    const bool in_img = rr >= 0 && rr < Hp;
    const float* __restrict__ src = h + ((int64_t)(b * Hp + (in_img ? rr : 0)) * kC + warp * 4) * W;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int j = lane + 32 * q, ww = w0 + j - HALO;
        pre[cc][q] = (in_img && j < COLS && ww >= 0 && ww < W) ? src[(int64_t)cc * W + ww] : 0.f;
      }
  };
  fetch(r0 - 1);
  for (int rr = r0 - 1; rr <= r_end; ++rr) {
    float* slot = ring + ((rr + 3) % 3) * kC * PITCH;
    const bool own = rr >= r0 && rr < r_end;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int c = warp * 4 + cc;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int j = lane + 32 * q;
        const float a = LEAKY ? leaky(pre[cc][q]) : pre[cc][q];
        if (j < COLS) slot[c * PITCH + j] = a;
        const int ww = w0 + j - HALO;
        if (NEXT != 0 && own && j >= HALO && j < HALO + kT_TW && ww < W) {
          if (NEXT == 1) next[((int64_t)(b * Hp + rr) * kC + c) * W + ww] = a;
          else next[((int64_t)(b * Hp + rr) * 2 * kC + 2 * c + (ww & 1)) * W2 + (ww >> 1)] = a;
        }
      }
    }
    if (NEXT == 2 && own && (W & 1) && w0 + kT_TW >= W && w0 < W && lane == 0) {
      // odd W: the last space-to-depth column of the odd phase has no source pixel
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        next[((int64_t)(b * Hp + rr) * 2 * kC + 2 * (warp * 4 + cc) + 1) * W2 + W2 - 1] = 0.f;
    }
    if (rr < r_end) fetch(rr + 1);
    __syncthreads();
    const int rs = rr - 1;  // output row (image coordinates) whose three input rows are now in the ring
    const bool do_row = rs >= r0 && rs < r_end && rs >= 1 && rs <= Hp - 2;
    if (do_row) {
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = warp * 4 + cc;
        float wk[WK];
#pragma unroll
        for (int q = 0; q < WK / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(wsm + c * WK + 4 * q);
          wk[4 * q] = v.x; wk[4 * q + 1] = v.y; wk[4 * q + 2] = v.z; wk[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int dr = 0; dr < 3; ++dr) {
          const float* row = ring + ((rs + dr - 1 + 3) % 3) * kC * PITCH + c * PITCH + 4 * lane;
          float v[4 + KW - 1 + 3];
#pragma unroll
          for (int q = 0; q < (4 + KW - 1 + 3) / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(row + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int dk = 0; dk < KW; ++dk)
#pragma unroll
            for (int p = 0; p < 4; ++p) a4[p] = fmaf(wk[dr * KW + dk], v[p + dk], a4[p]);
        }
      }
      *reinterpret_cast<float4*>(part + warp * kT_TW + 4 * lane) = make_float4(a4[0], a4[1], a4[2], a4[3]);
    }
    __syncthreads();
    if (do_row && tid < kT_TW) {
      float s = bias;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += part[q * kT_TW + tid];
      if (w0 + tid < W) score[((int64_t)b * bins + rs - 1) * W + w0 + tid] = s;
    }
  }
}
template <int KW>
constexpr size_t rows_to_map_smem() {
  return (size_t)(3 * kC * ((kT_TW + 2 * (KW / 2) + 3) & ~3) + kC * ((3 * KW + 3) & ~3) + 8 * kT_TW) * sizeof(float);
}

// ------------------------------------------------------------------------------------------------------ tail, backward
// dh[b,r,c,w] = leaky'(h) * ( sum_{dr,dk} wsc[c,dr,dk] * dscore[b, r - dr + 1 (image row), w - dk + 1] + d(next) )
// block: one image row, 256 columns, thread = column; the 9 dscore values of a column do not depend on the channel
constexpr int kB_TW = 256;
template <int NEXT>
__global__ void __launch_bounds__(kB_TW)
disc_tail_bwd_kernel(const float* __restrict__ h, const float* __restrict__ wsc, const float* __restrict__ dscore,
                     const float* __restrict__ dnext, float* __restrict__ dh, int Hp, int W) {
  __shared__ float wsm[kC * 9];
  __shared__ float dsm[3][kB_TW + 2];
  const int b = blockIdx.z, r = blockIdx.y, w0 = blockIdx.x * kB_TW, tid = threadIdx.x;
  const int bins = Hp - 2, W2 = (W + 1) >> 1, w = w0 + tid;
  float* __restrict__ out = dh + ((int64_t)(b * Hp + r) * kC) * W;
  if (r == 0 || r == Hp - 1) {
    if (w < W)
      for (int c = 0; c < kC; ++c) out[(int64_t)c * W + w] = 0.f;
    return;
  }
  for (int i = tid; i < kC * 9; i += kB_TW) wsm[i] = wsc[i];
  for (int i = tid; i < 3 * (kB_TW + 2); i += kB_TW) {
    const int j3 = i / (kB_TW + 2), j = i % (kB_TW + 2);
    const int rho = r - 1 + j3 - 1, ww = w0 + j - 1;  // score row (0-based) of image row r + j3 - 1
    dsm[j3][j] = (dscore && rho >= 0 && rho < bins && ww >= 0 && ww < W) ? dscore[((int64_t)b * bins + rho) * W + ww] : 0.f;
  }
  __syncthreads();
  if (w >= W) return;
  float d[9];  // d[dr*3+dk] = dscore[r - dr + 1, w - dk + 1]  ->  dsm[2 - dr][tid + 2 - dk]
#pragma unroll
  for (int dr = 0; dr < 3; ++dr)
#pragma unroll
    for (int dk = 0; dk < 3; ++dk) d[dr * 3 + dk] = dsm[2 - dr][tid + 2 - dk];
  const float* __restrict__ hr = h + ((int64_t)(b * Hp + r) * kC) * W;
  const float* __restrict__ nr = nullptr;
  if (NEXT == 1) nr = dnext + ((int64_t)(b * Hp + r) * kC) * W + w;
  if (NEXT == 2) nr = dnext + ((int64_t)(b * Hp + r) * 2 * kC + (w & 1)) * W2 + (w >> 1);
#pragma unroll 4
  for (int c = 0; c < kC; ++c) {
    float g = 0.f;
    if (NEXT == 1) g = nr[(int64_t)c * W];
    if (NEXT == 2) g = nr[(int64_t)c * 2 * W2];
#pragma unroll
    for (int k = 0; k < 9; ++k) g = fmaf(wsm[c * 9 + k], d[k], g);
    const float x = hr[(int64_t)c * W + w];
    out[(int64_t)c * W + w] = x > 0.f ? g : kSlope * g;
  }
}

// ---------------------------------------------------------------------------------------- score conv, weight gradient
// dw[c,dr,dk] += sum_{r,w} leaky(h[b, r + dr - 1, c, w + dk - 1]) * dscore[b, r, w];  db += sum dscore
// iterate over h elements (b, rr, c, w): a * dscore[rr - dr + 1, w - dk + 1].  persistent; warp = 4 channels,
// lane = column (4 x 32 columns per tile row); 36 accumulators per thread, one reduction + 289 atomics per block.
constexpr int kS_TW = 128, kS_Threads = 256;
__global__ void __launch_bounds__(kS_Threads)
disc_score_wgrad_kernel(const float* __restrict__ h, const float* __restrict__ dscore, float* __restrict__ dw,
                        float* __restrict__ db, int B, int Hp, int W) {
  __shared__ float dsm[3][kS_TW + 2];
  __shared__ float red[8][37];
  const int bins = Hp - 2, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wt = (W + kS_TW - 1) / kS_TW;
  const int64_t tiles = (int64_t)B * bins * wt;
  float acc[4][9];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[cc][k] = 0.f;
  float bsum = 0.f;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int xt = (int)(tile % wt);
    const int rr = (int)((tile / wt) % bins) + 1, b = (int)(tile / ((int64_t)wt * bins));  // interior image row of h
    const int w0 = xt * kS_TW;
    __syncthreads();
    for (int i = tid; i < 3 * (kS_TW + 2); i += kS_Threads) {
      const int j3 = i / (kS_TW + 2), j = i % (kS_TW + 2);
      const int rho = rr - 1 + j3 - 1, ww = w0 + j - 1;
      dsm[j3][j] = (rho >= 0 && rho < bins && ww >= 0 && ww < W) ? dscore[((int64_t)b * bins + rho) * W + ww] : 0.f;
    }
    __syncthreads();
    if (warp == 0) {  // bias gradient: the centre row of the dscore tile, interior columns
#pragma unroll
      for (int q = 0; q < 4; ++q) bsum += dsm[1][1 + lane + 32 * q];
    }
    const float* __restrict__ src = h + ((int64_t)(b * Hp + rr) * kC + warp * 4) * W;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = lane + 32 * q, w = w0 + j;
      float d[9];
#pragma unroll
      for (int dr = 0; dr < 3; ++dr)
#pragma unroll
        for (int dk = 0; dk < 3; ++dk) d[dr * 3 + dk] = dsm[2 - dr][j + 2 - dk];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float a = w < W ? leaky(src[(int64_t)cc * W + w]) : 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[cc][k] = fmaf(a, d[k], acc[cc][k]);
      }
    }
  }
#pragma unroll
  for (int cc = 0; cc < 4; ++cc)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      float v = acc[cc][k];
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[warp][cc * 9 + k] = v;
    }
  for (int off = 16; off; off >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, off);
  if (lane == 0) red[warp][36] = bsum;
  __syncthreads();
  for (int i = tid; i < 8 * 36; i += kS_Threads) {
    const int wq = i / 36, k = i % 36;  // channel = wq*4 + k/9
    atomicAdd(dw + (wq * 4 + k / 9) * 9 + k % 9, red[wq][k]);
  }
  if (tid == 0) atomicAdd(db, red[0][36]);
}

int persistent_grid() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * 2;
}

}  // namespace
}  // namespace sty

using namespace sty;

extern "C" int sty_disc_first_fwd(const float* y, const float* w, const float* bias, float* h, int B, int bins, int W,
                                  sty_stream_t stream) {
  STY_REQUIRE(y && w && bias && h && B > 0 && bins > 0 && W > 0 && bins + 2 <= 65535 && B <= 65535,
              "disc_first_fwd: bad argument");
  dim3 grid(cdiv(W, kF_TW), bins + 2, B);
  disc_first_fwd_kernel<<<grid, kF_Threads, 0, as_stream(stream)>>>(y, w, bias, h, bins, W);
  STY_CHECK_LAUNCH("disc_first_fwd");
  return STY_OK;
}

extern "C" int sty_disc_first_dgrad(const float* dh, const float* w, float* dy, int B, int bins, int W,
                                    sty_stream_t stream) {
  STY_REQUIRE(dh && w && dy && B > 0 && bins > 0 && W > 0 && B <= 65535, "disc_first_dgrad: bad argument");
  const size_t smem = rows_to_map_smem<9>();
  auto kern = disc_rows_to_map_kernel<0, 9, false, true>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(W, kT_TW), cdiv(bins + 2, kT_RS), B);
  kern<<<grid, kT_Threads, smem, as_stream(stream)>>>(dh, w, nullptr, dy, nullptr, bins + 2, W);
  STY_CHECK_LAUNCH("disc_first_dgrad");
  return STY_OK;
}

extern "C" int sty_disc_first_wgrad(const float* y, const float* dh, float* dw, float* db, int B, int bins, int W,
                                    sty_stream_t stream) {
  STY_REQUIRE(y && dh && dw && db && B > 0 && bins > 0 && W > 0, "disc_first_wgrad: bad argument");
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(dw, 0, kC * 27 * sizeof(float), st) != cudaSuccess ||
      cudaMemsetAsync(db, 0, kC * sizeof(float), st) != cudaSuccess) {
    set_error("disc_first_wgrad: memset failed");
    return STY_ERR_CUDA;
  }
  const size_t tile = (size_t)(kC * kW_Pitch + 3 * (kW_TW + 8)) * sizeof(float);
  const size_t red = (size_t)8 * kC * 28 * sizeof(float);
  const size_t smem = tile > red ? tile : red;
  cudaFuncSetAttribute(disc_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t tiles = (int64_t)B * bins * cdiv(W, kW_TW);
  const int grid = (int)(tiles < persistent_grid() ? tiles : persistent_grid());
  disc_first_wgrad_kernel<<<grid, kW_Threads, smem, st>>>(y, dh, dw, db, B, bins, W);
  STY_CHECK_LAUNCH("disc_first_wgrad");
  return STY_OK;
}

extern "C" int sty_disc_tail_fwd(const float* h, const float* w_score, const float* b_score, float* score, float* next,
                                 int next_kind, int B, int Hp, int W, sty_stream_t stream) {
  STY_REQUIRE(h && w_score && b_score && score && B > 0 && Hp > 2 && W > 0 && next_kind >= 0 && next_kind <= 2 &&
                  (next_kind == 0 || next) && B <= 65535,
              "disc_tail_fwd: bad argument");
  const size_t smem = rows_to_map_smem<3>();
  dim3 grid(cdiv(W, kT_TW), cdiv(Hp, kT_RS), B);
  cudaStream_t st = as_stream(stream);
#define STY_TAIL_FWD(N)                                                                              \
  {                                                                                                  \
    auto kern = disc_rows_to_map_kernel<N, 3, true, false>;                                          \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    kern<<<grid, kT_Threads, smem, st>>>(h, w_score, b_score, score, next, Hp, W);                   \
  }
  if (next_kind == 0) STY_TAIL_FWD(0)
  else if (next_kind == 1) STY_TAIL_FWD(1)
  else STY_TAIL_FWD(2)
#undef STY_TAIL_FWD
  STY_CHECK_LAUNCH("disc_tail_fwd");
  return STY_OK;
}

extern "C" int sty_disc_tail_bwd(const float* h, const float* w_score, const float* dscore, const float* dnext,
                                 float* dh, int next_kind, int B, int Hp, int W, sty_stream_t stream) {
  STY_REQUIRE(h && w_score && dh && B > 0 && Hp > 2 && W > 0 && next_kind >= 0 && next_kind <= 2 &&
                  (next_kind == 0 || dnext) && Hp <= 65535 && B <= 65535,
              "disc_tail_bwd: bad argument");
  dim3 grid(cdiv(W, kB_TW), Hp, B);
  cudaStream_t st = as_stream(stream);
  if (next_kind == 0) disc_tail_bwd_kernel<0><<<grid, kB_TW, 0, st>>>(h, w_score, dscore, dnext, dh, Hp, W);
  else if (next_kind == 1) disc_tail_bwd_kernel<1><<<grid, kB_TW, 0, st>>>(h, w_score, dscore, dnext, dh, Hp, W);
  else disc_tail_bwd_kernel<2><<<grid, kB_TW, 0, st>>>(h, w_score, dscore, dnext, dh, Hp, W);
  STY_CHECK_LAUNCH("disc_tail_bwd");
  return STY_OK;
}

extern "C" int sty_disc_score_wgrad(const float* h, const float* dscore, float* dw, float* db, int B, int Hp, int W,
                                    sty_stream_t stream) {
  STY_REQUIRE(h && dscore && dw && db && B > 0 && Hp > 2 && W > 0, "disc_score_wgrad: bad argument");
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(dw, 0, kC * 9 * sizeof(float), st) != cudaSuccess ||
      cudaMemsetAsync(db, 0, sizeof(float), st) != cudaSuccess) {
    set_error("disc_score_wgrad: memset failed");
    return STY_ERR_CUDA;
  }
  const int64_t tiles = (int64_t)B * (Hp - 2) * cdiv(W, kS_TW);
  const int grid = (int)(tiles < persistent_grid() * 2 ? tiles : persistent_grid() * 2);
  disc_score_wgrad_kernel<<<grid, kS_Threads, 0, st>>>(h, dscore, dw, db, B, Hp, W);
  STY_CHECK_LAUNCH("disc_score_wgrad");
  return STY_OK;
}
