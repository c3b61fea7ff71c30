// Harmonic source (SineGen + merge), conv-STFT of the excitation, and the
// spectral head + conv-iSTFT + tanh.  See include/stylish_b200.h for semantics.
//
// Numerics of the source: the reference accumulates phase = cumsum(f/sr)*2*pi*hop
// in fp32, where |phase| reaches 1e5..1e6 rad and one fp32 ulp is 0.01..0.1 rad,
// so its own output carries that rounding noise.  We carry the
// accumulated phase in fp64 *cycles* and reduce it mod 1 before the fp32 sin, so
// the excitation is closer to the exact-arithmetic value than the reference's
// fp32 evaluation; parity is judged against the fp64 oracle (tests/test_gpu_*.py).
#include <math.h>

#include "common.cuh"

namespace sty {

// linear-interpolation source coordinate of torch's F.interpolate(mode="linear",
// align_corners=False): src = (dst + 0.5) * (in/out) - 0.5, clamped at 0.
__device__ __forceinline__ void lerp_coord(double src, int n_in, int& i0, int& i1, double& lam) {
  if (src < 0.0) src = 0.0;
  i0 = (int)floor(src);
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = min(i0 + 1, n_in - 1);
  lam = src - (double)i0;
}

__device__ __forceinline__ double f0_up(const float* __restrict__ pitch,
                                        const float* __restrict__ voiced, int F, int hop, int j) {
  int i0, i1;
  double lam;
  lerp_coord(((double)j + 0.5) / (double)hop - 0.5, F, i0, i1, lam);
  const double a = (double)pitch[i0] * (double)voiced[i0];
  const double b = (double)pitch[i1] * (double)voiced[i1];
  return (1.0 - lam) * a + lam * b;
}

// one CTA per (b, harmonic): frame-rate phase increments computed in parallel, then an
// fp64 block scan (256 frames per pass, carry between passes)
__global__ void __launch_bounds__(256)
source_phase_kernel(const float* __restrict__ pitch, const float* __restrict__ voiced,
                    double* __restrict__ work, int B, int F, int hop, int H, double sr) {
  __shared__ double wsum[8];
  __shared__ double carry_s;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const float* __restrict__ pb = pitch + (int64_t)b * F;
  const float* __restrict__ vb = voiced + (int64_t)b * F;
  const int L = F * hop;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0.0;
  __syncthreads();
  for (int base = 0; base < F; base += 256) {
    const int i = base + tid;
    double v = 0.0;
    if (i < F) {
      int j0, j1;
      double lam;
      lerp_coord(((double)i + 0.5) * (double)hop - 0.5, L, j0, j1, lam);
      double r0 = f0_up(pb, vb, F, hop, j0) * (double)(h + 1) / sr;
      double r1 = f0_up(pb, vb, F, hop, j1) * (double)(h + 1) / sr;
      r0 -= floor(r0);
      r1 -= floor(r1);
      v = (1.0 - lam) * r0 + lam * r1;
    }
    // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    double off = carry_s;
    for (int k = 0; k < wid; ++k) off += wsum[k];
    v += off;
    if (i < F) work[((int64_t)b * H + h) * F + i] = v;
    __syncthreads();
    if (tid == 255) carry_s = v;
    __syncthreads();
  }
}

template <int HMAX>
__global__ void __launch_bounds__(256)
source_wave_kernel(const float* __restrict__ pitch, const float* __restrict__ voiced,
                   const float* __restrict__ noise, const float* __restrict__ lin_w,
                   const float* __restrict__ lin_b, const double* __restrict__ work,
                   float* __restrict__ out, int F, int hop, int H, float sine_amp, float noise_std,
                   float vthr) {
  const int b = blockIdx.y;
  const int L = F * hop;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L) return;
  const float* __restrict__ pb = pitch + (int64_t)b * F;
  const float* __restrict__ vb = voiced + (int64_t)b * F;
  int i0, i1;
  double lam;
  lerp_coord(((double)j + 0.5) / (double)hop - 0.5, F, i0, i1, lam);
  const double f0 = (1.0 - lam) * ((double)pb[i0] * (double)vb[i0]) + lam * ((double)pb[i1] * (double)vb[i1]);
  const float uv = ((float)f0 > vthr) ? 1.f : 0.f;
  const float namp = uv * noise_std + (1.f - uv) * sine_amp / 3.f;
  const float* __restrict__ nz = noise + ((int64_t)b * L + j) * H;
  float acc = lin_b[0];
#pragma unroll
  for (int h = 0; h < HMAX; ++h) {
    if (h < H) {
      const double* __restrict__ wr = work + ((int64_t)b * H + h) * F;
      double cyc = ((1.0 - lam) * wr[i0] + lam * wr[i1]) * (double)hop;
      cyc -= floor(cyc);
      const float sv = sinf((float)(cyc * 6.283185307179586476925286766559));
      const float val = sv * sine_amp * uv + namp * nz[h];
      acc = fmaf(lin_w[h], val, acc);
    }
  }
  out[(int64_t)b * L + j] = tanhf(acc);
}

// ---------------------------------------------------------------------- STFT
// thread = one frame; the NFFT-sample window sits in registers, the windowed DFT
// bases (bins_keep x NFFT, re and im) in shared memory (broadcast 128-bit loads).
template <int NFFT>
__global__ void __launch_bounds__(128)
stft_kernel(const float* __restrict__ wave, const float* __restrict__ basis_re,
            const float* __restrict__ basis_im, float* __restrict__ spec,
            float* __restrict__ phase, int64_t o_bs, int64_t o_cs, int L, int hop, int S, int bins_keep) {
  extern __shared__ __align__(16) float sm[];
  float* bre = sm;                      // [bins_keep][NFFT]
  float* bim = sm + bins_keep * NFFT;   // [bins_keep][NFFT]
  float* xs = bim + bins_keep * NFFT;   // [(128-1)*hop + NFFT]
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * 128;
  const int tid = threadIdx.x;
  for (int i = tid; i < bins_keep * NFFT; i += 128) {
    bre[i] = basis_re[i];
    bim[i] = basis_im[i];
  }
  const int span = 127 * hop + NFFT;
  const float* __restrict__ wb = wave + (int64_t)b * L;
  for (int i = tid; i < span; i += 128) {
    int n = f0 * hop + i - NFFT / 2;  // replicate padding (stft.py:105-107)
    n = max(0, min(L - 1, n));
    xs[i] = wb[n];
  }
  __syncthreads();
  const int f = f0 + tid;
  if (f >= S) return;
  float x[NFFT];
#pragma unroll
  for (int n = 0; n < NFFT; ++n) x[n] = xs[tid * hop + n];
  for (int kb = 0; kb < bins_keep; ++kb) {
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int n = 0; n < NFFT; n += 4) {
      const float4 cr = *reinterpret_cast<const float4*>(bre + kb * NFFT + n);
      const float4 ci = *reinterpret_cast<const float4*>(bim + kb * NFFT + n);
      re = fmaf(x[n], cr.x, re); re = fmaf(x[n + 1], cr.y, re);
      re = fmaf(x[n + 2], cr.z, re); re = fmaf(x[n + 3], cr.w, re);
      im = fmaf(x[n], ci.x, im); im = fmaf(x[n + 1], ci.y, im);
      im = fmaf(x[n + 2], ci.z, im); im = fmaf(x[n + 3], ci.w, im);
    }
    const float mag = sqrtf(re * re + im * im + 1e-14f);
    const int64_t o = (int64_t)b * o_bs + (int64_t)kb * o_cs + f;
    spec[o] = mag;
    phase[o] = atan2f(im / mag, re / mag);
  }
}

// ---------------------------------------------------------------- iSTFT head
// out[4 mo + r] = tanh( sum_{fr < 16} sum_{kb < 32}  re[kb, f] * Br[kb][4 (15 - fr) + r] - im[kb, f] * Bi[kb][...] ),
// f = mo - 7 + fr: a 16-tap, 64 -> 4 channel contraction per S-rate step (4096 FMAs per step, 3.9 GFMA per call at
// config 2).  Register-tiled: a thread owns G = 8 consecutive output groups (32 samples); per bin it loads its 23
// frames of re / im (12 LDS.128) and the 2 x 64 basis taps (32 broadcast LDS.128) for 1024 FMAs, so the loop is bound
// by the FMA pipe and not by shared-memory loads (the first version did 4 loads per 8 FMAs).  The head math
// (exp(logamp), cos / sin of atan2(imag, real) without the transcendental: one rsqrt) is done while staging.
template <int NFFT, int HOP, int BINS, int MT, int G, int KC>
__global__ void __launch_bounds__(MT)
istft_head_kernel(const float* __restrict__ logamp, int64_t logamp_bs, int64_t logamp_cs,
                  const float* __restrict__ real, const float* __restrict__ imag, int64_t ri_bs, int64_t ri_cs,
                  const float* __restrict__ basis_re, const float* __restrict__ basis_im,
                  float* __restrict__ out, int S) {
  static_assert(HOP == 4 && NFFT == 64 && G == 8, "register tiling assumes hop 4, n_fft 64, 8 groups per thread");
  constexpr int R = NFFT / HOP;          // 16 frames overlap each sample
  constexpr int FT = MT * G + R - 1;     // frames staged per block
  constexpr int FTP = (FT + 3) & ~3;
  constexpr int NF = G + R - 1;          // 23 frames per thread
  extern __shared__ __align__(16) float sm[];
  float* bre = sm;               // [KC][NFFT]  basis rows of the current chunk of bins
  float* bim = bre + KC * NFFT;  // [KC][NFFT]
  float* res = bim + KC * NFFT;  // [KC][FTP]
  float* ims = res + KC * FTP;   // [KC][FTP]
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int mo0 = blockIdx.x * MT * G;  // first output group of this CTA
  const int fbase = mo0 + R / 2 - (R - 1);
  float acc[G][HOP];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int r = 0; r < HOP; ++r) acc[g][r] = 0.f;

  for (int kb0 = 0; kb0 < BINS; kb0 += KC) {
    __syncthreads();  // previous chunk consumed
    for (int i = tid; i < KC * NFFT; i += MT) {
      bre[i] = basis_re[kb0 * NFFT + i];
      bim[i] = basis_im[kb0 * NFFT + i];
    }
    // staging: KC bins x FTP frames, several independent global loads in flight per thread
    const float* __restrict__ la = logamp + (int64_t)b * logamp_bs + (int64_t)kb0 * logamp_cs;
    const float* __restrict__ xr_ = real + (int64_t)b * ri_bs + (int64_t)kb0 * ri_cs;
    const float* __restrict__ yi_ = imag + (int64_t)b * ri_bs + (int64_t)kb0 * ri_cs;
    constexpr int NIT = (KC * FTP + MT - 1) / MT;
    constexpr int UN = (NIT % 6 == 0) ? 6 : (NIT % 5 == 0 ? 5 : 4);
#pragma unroll 1
    for (int it0 = 0; it0 < NIT; it0 += UN) {
      float lg[UN], vx[UN], vy[UN];
      bool ok[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = (it0 + u) * MT + tid;
        const int kc = i / FTP, fl = i - kc * FTP;
        const int f = fbase + fl;
        ok[u] = i < KC * FTP && fl < FT && f >= 0 && f <= S;
        const int fs = min(max(f, 0), S - 1);  // replicate-pad of one frame (generator.py:784-785)
        const int kcc = min(kc, KC - 1);
        lg[u] = ok[u] ? la[(int64_t)kcc * logamp_cs + fs] : 0.f;
        vx[u] = ok[u] ? xr_[(int64_t)kcc * ri_cs + fs] : 0.f;
        vy[u] = ok[u] ? yi_[(int64_t)kcc * ri_cs + fs] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int i = (it0 + u) * MT + tid;
        if (i >= KC * FTP) continue;
        float re = 0.f, im = 0.f;
        if (ok[u]) {
          const float mag = expf(lg[u]);
          // cos(atan2(y, x)) = x/|z|, sin(atan2(y, x)) = y/|z| (atan2(0, 0) = 0 -> cos 1, sin 0): one rsqrt
          // instead of atan2f + cosf + sinf; the operands are scaled first so x^2 + y^2 cannot over/underflow
          const float big = fmaxf(fabsf(vx[u]), fabsf(vy[u]));
          float c = 1.f, sn = 0.f;
          if (big > 0.f) {
            const float inv = 1.0f / big;
            const float xs = vx[u] * inv, ys = vy[u] * inv;
            const float rn = rsqrtf(fmaf(xs, xs, ys * ys));
            c = xs * rn;
            sn = ys * rn;
          }
          re = mag * c;
          im = mag * sn;
        }
        res[i] = re;      // i = kc * FTP + fl
        ims[i] = -im;     // sign folded here: acc += re * Br + (-im) * Bi
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int kc = 0; kc < KC; ++kc) {
      float xr[NF + 1], xi[NF + 1];
      const float4* rp = reinterpret_cast<const float4*>(res + kc * FTP + tid * G);
      const float4* ip = reinterpret_cast<const float4*>(ims + kc * FTP + tid * G);
#pragma unroll
      for (int q = 0; q < (NF + 1) / 4; ++q) {
        const float4 a = rp[q], c = ip[q];
        xr[4 * q] = a.x; xr[4 * q + 1] = a.y; xr[4 * q + 2] = a.z; xr[4 * q + 3] = a.w;
        xi[4 * q] = c.x; xi[4 * q + 1] = c.y; xi[4 * q + 2] = c.z; xi[4 * q + 3] = c.w;
      }
      const float4* cr4 = reinterpret_cast<const float4*>(bre + kc * NFFT);
      const float4* ci4 = reinterpret_cast<const float4*>(bim + kc * NFFT);
#pragma unroll
      for (int fr = 0; fr < R; ++fr) {
        const float4 cr = cr4[R - 1 - fr], ci = ci4[R - 1 - fr];  // taps 4 (15 - fr) .. + 3
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float re = xr[g + fr], im = xi[g + fr];
          acc[g][0] = fmaf(re, cr.x, acc[g][0]); acc[g][0] = fmaf(im, ci.x, acc[g][0]);
          acc[g][1] = fmaf(re, cr.y, acc[g][1]); acc[g][1] = fmaf(im, ci.y, acc[g][1]);
          acc[g][2] = fmaf(re, cr.z, acc[g][2]); acc[g][2] = fmaf(im, ci.z, acc[g][2]);
          acc[g][3] = fmaf(re, cr.w, acc[g][3]); acc[g][3] = fmaf(im, ci.w, acc[g][3]);
        }
      }
    }
  }
  float* __restrict__ ob = out + (int64_t)b * S * HOP;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int mo = mo0 + tid * G + g;
    if (mo < S)
      *reinterpret_cast<float4*>(ob + (int64_t)mo * HOP) =
          make_float4(tanhf(acc[g][0]), tanhf(acc[g][1]), tanhf(acc[g][2]), tanhf(acc[g][3]));
  }
}

}  // namespace sty

using namespace sty;

extern "C" int sty_source_fwd(const float* pitch, const float* voiced, const float* noise,
                              const float* lin_w, const float* lin_b, double* work, float* out,
                              int B, int F, int hop, int H, float sample_rate, float sine_amp,
                              float noise_std, float voiced_threshold, sty_stream_t stream) {
  STY_REQUIRE(pitch && voiced && noise && lin_w && lin_b && work && out, "source: null pointer");
  STY_REQUIRE(B > 0 && F > 1 && hop > 0 && H > 0 && H <= 16, "source: bad shape (H<=16)");
  cudaStream_t st = as_stream(stream);
  source_phase_kernel<<<B * H, 256, 0, st>>>(pitch, voiced, work, B, F, hop, H, (double)sample_rate);
  STY_CHECK_LAUNCH("source_phase");
  dim3 grid(cdiv((int64_t)F * hop, 256), B);
  source_wave_kernel<16><<<grid, 256, 0, st>>>(pitch, voiced, noise, lin_w, lin_b, work, out, F, hop,
                                               H, sine_amp, noise_std, voiced_threshold);
  STY_CHECK_LAUNCH("source_wave");
  return STY_OK;
}

extern "C" int sty_stft_pitched_fwd(const float* wave, const float* basis_re, const float* basis_im,
                                    float* spec, float* phase, int64_t out_bs, int64_t out_cs, int B, int L,
                                    int n_fft, int hop, int bins_keep, sty_stream_t stream) {
  STY_REQUIRE(wave && basis_re && basis_im && spec && phase, "stft: null pointer");
  STY_REQUIRE(n_fft == 64, "stft: built for n_fft=64 (got %d)", n_fft);
  STY_REQUIRE(B > 0 && L > 0 && hop > 0 && L % hop == 0 && bins_keep > 0 && bins_keep <= n_fft / 2 + 1,
              "stft: bad shape");
  const int S = L / hop;  // frames kept (the trailing frame is dropped, generator.py:725,728)
  STY_REQUIRE(out_cs >= S && out_bs >= out_cs * bins_keep, "stft: output strides too small");
  const size_t smem = ((size_t)2 * bins_keep * n_fft + 127 * hop + n_fft) * sizeof(float);
  STY_REQUIRE(smem <= 48 * 1024, "stft: hop too large for the staging buffer");
  dim3 grid(cdiv(S, 128), B);
  stft_kernel<64><<<grid, 128, smem, as_stream(stream)>>>(wave, basis_re, basis_im, spec, phase, out_bs, out_cs,
                                                         L, hop, S, bins_keep);
  STY_CHECK_LAUNCH("stft");
  return STY_OK;
}

extern "C" int sty_stft_fwd(const float* wave, const float* basis_re, const float* basis_im,
                            float* spec, float* phase, int B, int L, int n_fft, int hop,
                            int bins_keep, sty_stream_t stream) {
  const int64_t S = hop > 0 ? L / hop : 0;
  return sty_stft_pitched_fwd(wave, basis_re, basis_im, spec, phase, S * bins_keep, S, B, L, n_fft, hop,
                              bins_keep, stream);
}

extern "C" int sty_istft_head_pitched_fwd(const float* logamp, int64_t logamp_bs, int64_t logamp_cs,
                                          const float* real, const float* imag, int64_t ri_bs, int64_t ri_cs,
                                          const float* basis_re, const float* basis_im, float* out, int B, int S,
                                          int bins, int n_fft, int hop, sty_stream_t stream) {
  STY_REQUIRE(logamp && real && imag && basis_re && basis_im && out, "istft_head: null pointer");
  STY_REQUIRE(n_fft == 64 && hop == 4 && bins == 32,
              "istft_head: built for n_fft=64 hop=4 bins=32 (got %d %d %d)", n_fft, hop, bins);
  STY_REQUIRE(B > 0 && S > 0 && logamp_cs >= S && ri_cs >= S, "istft_head: bad shape");
  constexpr int MT = 128, G = 8, KC = 2, R = 16, FTP = (MT * G + R - 1 + 3) & ~3;  // 17.7 KB: 7 CTAs per SM
  const size_t smem = ((size_t)2 * KC * 64 + 2 * KC * FTP) * sizeof(float);
  auto kern = istft_head_kernel<64, 4, 32, MT, G, KC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid(cdiv(S, MT * G), B);
  kern<<<grid, MT, smem, as_stream(stream)>>>(logamp, logamp_bs, logamp_cs, real, imag, ri_bs, ri_cs, basis_re,
                                              basis_im, out, S);
  STY_CHECK_LAUNCH("istft_head");
  return STY_OK;
}

extern "C" int sty_istft_head_fwd(const float* logamp, int64_t logamp_bs, const float* real,
                                  const float* imag, int64_t ri_bs, const float* basis_re,
                                  const float* basis_im, float* out, int B, int S, int bins,
                                  int n_fft, int hop, sty_stream_t stream) {
  return sty_istft_head_pitched_fwd(logamp, logamp_bs, S, real, imag, ri_bs, S, basis_re, basis_im, out, B, S,
                                    bins, n_fft, hop, stream);
}
