// Image-side kernels of the mel style encoder (mel_style_encoder.py): everything that is not a
// stride-1 convolution.  Images are kept "row-channel": X[b, r, c, w] contiguous as (B, Hp, C, W) with
// Hp = H + 2 and all-zero border rows r = 0 and r = Hp-1, so that a 3x3 (or 5x5) Conv2d is the stride-1
// Conv1d kernel applied to R consecutive rows seen as R*C stacked channels (sty_conv1d_fwd on a strided,
// overlapping view) — the tensor-core path, its data gradient and its weight gradient are reused as they are.
// Here: the fold that turns the gradient of that overlapping view back into an image, the learned
// depthwise stride-2 3x3 downsampling conv, the 2x2 average pool, and the region mean.
#include "common.cuh"

namespace sty {
namespace {

// dx[b,r,c,w] = sum_{k<R} g[n = b*Hp + r - k][k*C + c][w]   for 0 <= n < N
__global__ void __launch_bounds__(256)
fold_rows_kernel(const float* __restrict__ g, float* __restrict__ dx, int R, int Hp, int C, int W, int64_t N) {
  const int64_t row = blockIdx.x;  // b*Hp + r
  const int c = blockIdx.y;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < R; ++k) {
      const int64_t n = row - k;
      if (n >= 0 && n < N) a += g[(n * R * C + (int64_t)k * C + c) * W + w];
    }
    dx[(row * C + c) * W + w] = a;
  }
}

// depthwise 3x3, stride 2, padding 1 on row-channel images (padded row index = image row + 1)
__global__ void __launch_bounds__(128)
dw3x3s2_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, int Hp, int C, int W, int Hop, int Wo) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int ro = blockIdx.x;  // output padded row
  float* __restrict__ yr = y + (((int64_t)b * Hop + ro) * C + c) * Wo;
  if (ro == 0 || ro == Hop - 1) {
    for (int j = threadIdx.x; j < Wo; j += blockDim.x) yr[j] = 0.f;
    return;
  }
  const int i = ro - 1;
  float wk[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wk[k] = w[c * 9 + k];
  const float bv = bias ? bias[c] : 0.f;
  for (int j = threadIdx.x; j < Wo; j += blockDim.x) {
    float a = bv;
#pragma unroll
    for (int aa = 0; aa < 3; ++aa) {
      const int r = 2 * i + aa;  // padded input row (image row 2i+aa-1); rows 0 and Hp-1 are zero
      if (r >= Hp) continue;
      const float* __restrict__ xr = x + (((int64_t)b * Hp + r) * C + c) * W;
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const int col = 2 * j + bb - 1;
        if (col >= 0 && col < W) a = fmaf(wk[aa * 3 + bb], xr[col], a);
      }
    }
    yr[j] = a;
  }
}

// dx of the above (border rows written as zero)
__global__ void __launch_bounds__(128)
dw3x3s2_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int Hp,
                      int C, int W, int Hop, int Wo) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int r = blockIdx.x;  // padded input row
  float* __restrict__ dr = dx + (((int64_t)b * Hp + r) * C + c) * W;
  if (r == 0 || r == Hp - 1) {
    for (int col = threadIdx.x; col < W; col += blockDim.x) dr[col] = 0.f;
    return;
  }
  float wk[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) wk[k] = w[c * 9 + k];
  for (int col = threadIdx.x; col < W; col += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int aa = 0; aa < 3; ++aa) {
      const int t = r - aa;  // = 2i
      if (t < 0 || (t & 1)) continue;
      const int i = t >> 1;
      if (i >= Hop - 2) continue;
      const float* __restrict__ gr = dy + (((int64_t)b * Hop + i + 1) * C + c) * Wo;
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const int u = col + 1 - bb;  // = 2j
        if (u < 0 || (u & 1)) continue;
        const int j = u >> 1;
        if (j < Wo) a = fmaf(wk[aa * 3 + bb], gr[j], a);
      }
    }
    dr[col] = a;
  }
}

// dw[c, 9] and db[c] of the above: one CTA per (c, b)
__global__ void __launch_bounds__(256)
dw3x3s2_bwd_w_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                     float* __restrict__ db, int Hp, int C, int W, int Hop, int Wo) {
  __shared__ float red[32];
  const int c = blockIdx.x, b = blockIdx.y;
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.f;
  const int Ho = Hop - 2;
  for (int idx = threadIdx.x; idx < Ho * Wo; idx += blockDim.x) {
    const int i = idx / Wo, j = idx - i * Wo;
    const float g = dy[(((int64_t)b * Hop + i + 1) * C + c) * Wo + j];
    acc[9] += g;
#pragma unroll
    for (int aa = 0; aa < 3; ++aa) {
      const int r = 2 * i + aa;
      if (r >= Hp) continue;
      const float* __restrict__ xr = x + (((int64_t)b * Hp + r) * C + c) * W;
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) {
        const int col = 2 * j + bb - 1;
        if (col >= 0 && col < W) acc[aa * 3 + bb] = fmaf(g, xr[col], acc[aa * 3 + bb]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float s = block_sum(acc[k], red);
    if (threadIdx.x == 0) {
      if (k < 9) atomicAdd(dw + c * 9 + k, s);
      else if (db) atomicAdd(db + c, s);
    }
  }
}

// 2x2 average pool (odd W: the last column is replicated first, mel_style_encoder.py:57-60)
__global__ void __launch_bounds__(128)
avgpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int Hp, int C, int W, int Hop, int Wo) {
  const int c = blockIdx.y, b = blockIdx.z, ro = blockIdx.x;
  float* __restrict__ yr = y + (((int64_t)b * Hop + ro) * C + c) * Wo;
  if (ro == 0 || ro == Hop - 1) {
    for (int j = threadIdx.x; j < Wo; j += blockDim.x) yr[j] = 0.f;
    return;
  }
  const float* __restrict__ x0 = x + (((int64_t)b * Hp + 2 * ro - 1) * C + c) * W;
  const float* __restrict__ x1 = x0 + (int64_t)C * W;
  for (int j = threadIdx.x; j < Wo; j += blockDim.x) {
    const int a = 2 * j, bcol = min(2 * j + 1, W - 1);
    yr[j] = 0.25f * ((x0[a] + x0[bcol]) + (x1[a] + x1[bcol]));
  }
}

__global__ void __launch_bounds__(128)
avgpool2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int Hp, int C, int W, int Hop, int Wo) {
  const int c = blockIdx.y, b = blockIdx.z, r = blockIdx.x;
  float* __restrict__ dr = dx + (((int64_t)b * Hp + r) * C + c) * W;
  if (r == 0 || r == Hp - 1 || (r - 1) / 2 >= Hop - 2) {
    for (int col = threadIdx.x; col < W; col += blockDim.x) dr[col] = 0.f;
    return;
  }
  const float* __restrict__ gr = dy + (((int64_t)b * Hop + (r - 1) / 2 + 1) * C + c) * Wo;
  for (int col = threadIdx.x; col < W; col += blockDim.x) {
    float v = 0.25f * gr[col >> 1];
    if ((W & 1) && col == W - 1) v *= 2.f;  // replicated column: read twice by the last window
    dr[col] = v;
  }
}

// out[b,c] = mean over rows [r0,r0+Rn) x cols [w0,w0+Wn) of x (B,Hp,C,W)
__global__ void __launch_bounds__(256)
region_mean_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int Hp, int C, int W, int r0, int Rn,
                       int w0, int Wn) {
  __shared__ float red[32];
  const int c = blockIdx.x, b = blockIdx.y;
  float s = 0.f;
  for (int idx = threadIdx.x; idx < Rn * Wn; idx += blockDim.x) {
    const int r = idx / Wn, w = idx - r * Wn;
    s += x[(((int64_t)b * Hp + r0 + r) * C + c) * W + w0 + w];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[b * C + c] = s / (float)(Rn * Wn);
}

__global__ void __launch_bounds__(128)
region_mean_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int Hp, int C, int W, int r0, int Rn,
                       int w0, int Wn) {
  const int c = blockIdx.y, b = blockIdx.z, r = blockIdx.x;
  float* __restrict__ dr = dx + (((int64_t)b * Hp + r) * C + c) * W;
  const bool in_r = r >= r0 && r < r0 + Rn;
  const float v = g[b * C + c] / (float)(Rn * Wn);
  for (int col = threadIdx.x; col < W; col += blockDim.x)
    dr[col] = (in_r && col >= w0 && col < w0 + Wn) ? v : 0.f;
}

}  // namespace
}  // namespace sty

using namespace sty;

extern "C" int sty_fold_rows(const float* g, float* dx, int R, int B, int Hp, int C, int W, sty_stream_t stream) {
  STY_REQUIRE(g && dx && R >= 1 && B > 0 && Hp >= R && C > 0 && C <= 65535 && W > 0, "fold_rows: bad argument");
  const int64_t N = (int64_t)B * Hp - (R - 1);
  dim3 grid((unsigned)((int64_t)B * Hp), C);
  fold_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, dx, R, Hp, C, W, N);
  STY_CHECK_LAUNCH("fold_rows");
  return STY_OK;
}

extern "C" int sty_dwconv3x3s2_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Hp, int C,
                                   int W, sty_stream_t stream) {
  STY_REQUIRE(x && w && y && B > 0 && Hp > 2 && C > 0 && C <= 65535 && W > 0 && B <= 65535, "dwconv3x3s2: bad argument");
  const int Hop = (Hp - 2 - 1) / 2 + 1 + 2, Wo = (W - 1) / 2 + 1;
  dim3 grid(Hop, C, B);
  dw3x3s2_fwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, w, bias, y, Hp, C, W, Hop, Wo);
  STY_CHECK_LAUNCH("dwconv3x3s2_fwd");
  return STY_OK;
}

extern "C" int sty_dwconv3x3s2_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db,
                                   int B, int Hp, int C, int W, sty_stream_t stream) {
  STY_REQUIRE(dy && x && w && dw && B > 0 && Hp > 2 && C > 0 && C <= 65535 && W > 0 && B <= 65535,
              "dwconv3x3s2_bwd: bad argument");
  const int Hop = (Hp - 2 - 1) / 2 + 1 + 2, Wo = (W - 1) / 2 + 1;
  cudaStream_t st = as_stream(stream);
  if (dx) {
    dim3 grid(Hp, C, B);
    dw3x3s2_bwd_dx_kernel<<<grid, 128, 0, st>>>(dy, w, dx, Hp, C, W, Hop, Wo);
    STY_CHECK_LAUNCH("dwconv3x3s2_bwd_dx");
  }
  dim3 gw(C, B);
  dw3x3s2_bwd_w_kernel<<<gw, 256, 0, st>>>(dy, x, dw, db, Hp, C, W, Hop, Wo);
  STY_CHECK_LAUNCH("dwconv3x3s2_bwd_w");
  return STY_OK;
}

extern "C" int sty_avgpool2_fwd(const float* x, float* y, int B, int Hp, int C, int W, sty_stream_t stream) {
  STY_REQUIRE(x && y && B > 0 && Hp > 2 && (Hp - 2) % 2 == 0 && C > 0 && C <= 65535 && W > 0 && B <= 65535,
              "avgpool2: bad argument (H must be even)");
  const int Hop = (Hp - 2) / 2 + 2, Wo = (W + 1) / 2;
  dim3 grid(Hop, C, B);
  avgpool2_fwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(x, y, Hp, C, W, Hop, Wo);
  STY_CHECK_LAUNCH("avgpool2_fwd");
  return STY_OK;
}

extern "C" int sty_avgpool2_bwd(const float* dy, float* dx, int B, int Hp, int C, int W, sty_stream_t stream) {
  STY_REQUIRE(dy && dx && B > 0 && Hp > 2 && (Hp - 2) % 2 == 0 && C > 0 && C <= 65535 && W > 0 && B <= 65535,
              "avgpool2_bwd: bad argument");
  const int Hop = (Hp - 2) / 2 + 2, Wo = (W + 1) / 2;
  dim3 grid(Hp, C, B);
  avgpool2_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(dy, dx, Hp, C, W, Hop, Wo);
  STY_CHECK_LAUNCH("avgpool2_bwd");
  return STY_OK;
}

extern "C" int sty_region_mean_fwd(const float* x, float* out, int B, int Hp, int C, int W, int r0, int Rn, int w0,
                                   int Wn, sty_stream_t stream) {
  STY_REQUIRE(x && out && B > 0 && C > 0 && r0 >= 0 && Rn > 0 && r0 + Rn <= Hp && w0 >= 0 && Wn > 0 && w0 + Wn <= W,
              "region_mean: bad argument");
  dim3 grid(C, B);
  region_mean_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, out, Hp, C, W, r0, Rn, w0, Wn);
  STY_CHECK_LAUNCH("region_mean_fwd");
  return STY_OK;
}

extern "C" int sty_region_mean_bwd(const float* g, float* dx, int B, int Hp, int C, int W, int r0, int Rn, int w0,
                                   int Wn, sty_stream_t stream) {
  STY_REQUIRE(g && dx && B > 0 && C > 0 && C <= 65535 && B <= 65535 && r0 >= 0 && Rn > 0 && r0 + Rn <= Hp && w0 >= 0 &&
                  Wn > 0 && w0 + Wn <= W, "region_mean_bwd: bad argument");
  dim3 grid(Hp, C, B);
  region_mean_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(g, dx, Hp, C, W, r0, Rn, w0, Wn);
  STY_CHECK_LAUNCH("region_mean_bwd");
  return STY_OK;
}
