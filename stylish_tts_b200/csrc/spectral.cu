// Framed real-FFT front-end fused with magnitude / phase / mel filterbank / log
// (forward and backward), plus the STFT-loss reductions.  See
// include/stylish_b200.h for semantics and the reference lines each entry replaces.
//
// Design: one CTA owns 2*PAIRS consecutive frames of one signal.  Two real frames
// are packed as the real/imaginary parts of ONE complex radix-2 FFT in shared
// memory (X1 = (Z[k]+conj Z[N-k])/2, X2 = (Z[k]-conj Z[N-k])/2i), so a CTA runs
// PAIRS complex FFTs side by side (128 threads each).  The forward transform is
// decimation-in-time (bit-reversed placement at load, natural-order spectrum);
// the backward transform is decimation-in-frequency on the same buffer (natural
// order in, bit-reversed out), so the gradient spectrum is built IN PLACE over the
// recomputed forward spectrum and no second buffer is needed.  Magnitudes are
// staged frame-minor in shared memory so that the (B, K, N) outputs are written
// as 32-byte runs and the sparse (triangular) mel filterbank is applied from
// shared memory.  Everything between the audio samples and the mel / phase /
// magnitude tensors stays on chip.
#include <math.h>

#include "common.cuh"

namespace sty {

namespace {

constexpr int GROUP = 128;  // threads per complex FFT

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// in-place radix-2 DIT: input bit-reversed, output natural.  tw[q] = exp(-2 pi i q / N).
// All threads of the CTA call this together (__syncthreads inside).
__device__ __forceinline__ void fft_dit(float2* z, const float2* tw, int logn, int gt) {
  const int n2 = 1 << (logn - 1);
  for (int s = 1; s <= logn; ++s) {
    const int half = 1 << (s - 1);
    const int tstep = n2 >> (s - 1);
    for (int j = gt; j < n2; j += GROUP) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos;
      const int i1 = i0 + half;
      const float2 a = z[i0];
      const float2 t = cmul(tw[pos * tstep], z[i1]);
      z[i0] = make_float2(a.x + t.x, a.y + t.y);
      z[i1] = make_float2(a.x - t.x, a.y - t.y);
    }
    __syncthreads();
  }
}

// in-place radix-2 DIF: input natural, output bit-reversed.
__device__ __forceinline__ void fft_dif(float2* z, const float2* tw, int logn, int gt) {
  const int n2 = 1 << (logn - 1);
  for (int s = logn; s >= 1; --s) {
    const int half = 1 << (s - 1);
    const int tstep = n2 >> (s - 1);
    for (int j = gt; j < n2; j += GROUP) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos;
      const int i1 = i0 + half;
      const float2 a = z[i0], b = z[i1];
      z[i0] = make_float2(a.x + b.x, a.y + b.y);
      z[i1] = cmul(tw[pos * tstep], make_float2(a.x - b.x, a.y - b.y));
    }
    __syncthreads();
  }
}

__device__ __forceinline__ int reflect_index(int i, int L) {
  if (i < 0) i = -i;
  if (i >= L) i = 2 * (L - 1) - i;
  return i;
}

// spectra of the two real frames packed in one complex FFT
__device__ __forceinline__ void unpack_pair(const float2* z, int k, int N, float2& x1, float2& x2) {
  if (k == 0 || k == N / 2) {
    const float2 a = z[k];
    x1 = make_float2(a.x, 0.f);
    x2 = make_float2(a.y, 0.f);
  } else {
    const float2 a = z[k], c = z[N - k];
    x1 = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
    x2 = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));
  }
}

__device__ __forceinline__ float mel_post(float raw, int mode, float eps, float mean, float inv_std) {
  if (mode == 1) return log1pf(raw);
  if (mode == 2) return (logf(eps + raw) - mean) * inv_std;
  return raw;
}

struct SpecShared {
  float2* z;     // [PAIRS][N]
  float2* tw;    // [N/2]
  float* magS;   // [K][FR]   (FR = 2*PAIRS frames, frame-minor)
  float* melS;   // [n_mels][FR]
};

// load + window + FFT of the CTA's frames; leaves the packed spectra in sh.z
template <int PAIRS>
__device__ __forceinline__ void load_and_fft(const sty_spectrogram_args& a, const SpecShared& sh,
                                             int b, int f0, int logn) {
  const int N = a.n_fft;
  const int tid = threadIdx.x;
  const float* __restrict__ ab = a.audio + (int64_t)b * a.audio_bs;
  for (int i = tid; i < N / 2; i += blockDim.x) sh.tw[i] = reinterpret_cast<const float2*>(a.twiddle)[i];
  for (int idx = tid; idx < PAIRS * N; idx += blockDim.x) {
    const int p = idx / N, n = idx - p * N;
    const int fa = f0 + 2 * p, fb = fa + 1;
    const float w = a.window[n];
    float va = 0.f, vb = 0.f;
    if (fa < a.n_frames) va = w * ab[reflect_index(fa * a.hop + n - N / 2, a.L)];
    if (fb < a.n_frames) vb = w * ab[reflect_index(fb * a.hop + n - N / 2, a.L)];
    const int r = (int)(__brev((unsigned)n) >> (32 - logn));
    sh.z[p * N + r] = make_float2(va, vb);
  }
  __syncthreads();
  const int grp = tid / GROUP, gt = tid - grp * GROUP;
  fft_dit(sh.z + grp * N, sh.tw, logn, gt);
}

// magnitudes (power 1 or 2) into magS, then the sparse mel filterbank into melS (raw)
template <int PAIRS>
__device__ __forceinline__ void mag_and_mel(const sty_spectrogram_args& a, const SpecShared& sh) {
  constexpr int FR = 2 * PAIRS;
  const int N = a.n_fft, K = N / 2 + 1;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < PAIRS * K; idx += blockDim.x) {
    const int p = idx / K, k = idx - p * K;
    float2 x1, x2;
    unpack_pair(sh.z + p * N, k, N, x1, x2);
    float m1 = x1.x * x1.x + x1.y * x1.y, m2 = x2.x * x2.x + x2.y * x2.y;
    if (a.power == 1) {
      m1 = sqrtf(m1);
      m2 = sqrtf(m2);
    }
    sh.magS[k * FR + 2 * p] = m1;
    sh.magS[k * FR + 2 * p + 1] = m2;
  }
  __syncthreads();
  if (a.n_mels > 0) {
    for (int idx = tid; idx < a.n_mels * FR; idx += blockDim.x) {
      const int m = idx / FR, fr = idx - m * FR;
      const int s = a.fb_start[m], len = a.fb_len[m];
      const float* __restrict__ w = a.fb_w + a.fb_off[m];
      float acc = 0.f;
      for (int j = 0; j < len; ++j) acc = fmaf(w[j], sh.magS[(s + j) * FR + fr], acc);
      sh.melS[idx] = acc;
    }
    __syncthreads();
  }
}

template <int PAIRS>
__device__ __forceinline__ SpecShared carve(float* sm, int N, int n_mels) {
  constexpr int FR = 2 * PAIRS;
  SpecShared sh;
  sh.z = reinterpret_cast<float2*>(sm);
  sh.tw = sh.z + PAIRS * N;
  sh.magS = reinterpret_cast<float*>(sh.tw + N / 2);
  sh.melS = sh.magS + (N / 2 + 1) * FR;
  (void)n_mels;
  return sh;
}

template <int PAIRS>
static size_t spec_smem_bytes(int N, int n_mels) {
  constexpr int FR = 2 * PAIRS;
  return (size_t)PAIRS * N * 8 + (size_t)(N / 2) * 8 + (size_t)(N / 2 + 1) * FR * 4 +
         (size_t)(n_mels > 0 ? n_mels : 1) * FR * 4 + 64;
}

template <int PAIRS>
__global__ void __launch_bounds__(PAIRS* GROUP)
spectrogram_fwd_kernel(const sty_spectrogram_args a, int logn) {
  extern __shared__ __align__(16) float sm[];
  constexpr int FR = 2 * PAIRS;
  const int N = a.n_fft, K = N / 2 + 1;
  const SpecShared sh = carve<PAIRS>(sm, N, a.n_mels);
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FR;
  const int tid = threadIdx.x;
  load_and_fft<PAIRS>(a, sh, b, f0, logn);
  // phase straight from the packed spectra (before the buffer is reused)
  if (a.phase) {
    float* __restrict__ ph = a.phase + (int64_t)b * K * a.n_frames;
    for (int idx = tid; idx < K * PAIRS; idx += blockDim.x) {
      const int k = idx / PAIRS, p = idx - k * PAIRS;
      float2 x1, x2;
      unpack_pair(sh.z + p * N, k, N, x1, x2);
      const int fa = f0 + 2 * p;
      const float m1 = sqrtf(x1.x * x1.x + x1.y * x1.y), m2 = sqrtf(x2.x * x2.x + x2.y * x2.y);
      if (fa < a.n_frames) ph[(int64_t)k * a.n_frames + fa] = m1 > a.phase_floor ? atan2f(x1.y, x1.x) : 0.f;
      if (fa + 1 < a.n_frames)
        ph[(int64_t)k * a.n_frames + fa + 1] = m2 > a.phase_floor ? atan2f(x2.y, x2.x) : 0.f;
    }
  }
  mag_and_mel<PAIRS>(a, sh);
  if (a.mag) {
    float* __restrict__ mg = a.mag + (int64_t)b * K * a.n_frames;
    for (int idx = tid; idx < K * FR; idx += blockDim.x) {
      const int k = idx / FR, fr = idx - k * FR;
      if (f0 + fr < a.n_frames) mg[(int64_t)k * a.n_frames + f0 + fr] = sh.magS[idx];
    }
  }
  if (a.mel) {
    float* __restrict__ ml = a.mel + (int64_t)b * a.n_mels * a.n_frames;
    const float inv_std = 1.f / a.mel_std;
    for (int idx = tid; idx < a.n_mels * FR; idx += blockDim.x) {
      const int m = idx / FR, fr = idx - m * FR;
      if (f0 + fr < a.n_frames)
        ml[(int64_t)m * a.n_frames + f0 + fr] = mel_post(sh.melS[idx], a.mel_mode, a.mel_eps, a.mel_mean, inv_std);
    }
  }
}

// Backward: d_audio += d(mel, phase, mag)/d(audio).  Recomputes the forward spectrum.
template <int PAIRS>
__global__ void __launch_bounds__(PAIRS* GROUP)
spectrogram_bwd_kernel(const sty_spectrogram_args a, const float* __restrict__ d_mel,
                       const float* __restrict__ d_phase, const float* __restrict__ d_mag,
                       float* __restrict__ d_audio, int64_t d_audio_bs, int logn) {
  extern __shared__ __align__(16) float sm[];
  constexpr int FR = 2 * PAIRS;
  const int N = a.n_fft, K = N / 2 + 1;
  const SpecShared sh = carve<PAIRS>(sm, N, a.n_mels);
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FR;
  const int tid = threadIdx.x;
  load_and_fft<PAIRS>(a, sh, b, f0, logn);
  mag_and_mel<PAIRS>(a, sh);
  // melS <- d loss / d raw mel
  if (a.n_mels > 0) {
    const float inv_std = 1.f / a.mel_std;
    for (int idx = tid; idx < a.n_mels * FR; idx += blockDim.x) {
      const int m = idx / FR, fr = idx - m * FR;
      float g = 0.f;
      if (d_mel && f0 + fr < a.n_frames) {
        g = d_mel[((int64_t)b * a.n_mels + m) * a.n_frames + f0 + fr];
        const float raw = sh.melS[idx];
        if (a.mel_mode == 1) g = g / (1.f + raw);
        else if (a.mel_mode == 2) g = g * inv_std / (a.mel_eps + raw);
      }
      sh.melS[idx] = g;
    }
    __syncthreads();
  }
  // gradient spectrum, in place: slot k <- conj(H[k]), slot N-k <- conj(H[N-k]) where
  // H = G1' + i G2', G' the Hermitian extension of G = dRe + i dIm (halved off the axis)
  for (int idx = tid; idx < PAIRS * K; idx += blockDim.x) {
    const int p = idx / K, k = idx - p * K;
    float2* z = sh.z + p * N;
    float2 x[2];
    unpack_pair(z, k, N, x[0], x[1]);
    float2 G[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int fr = 2 * p + q;
      const bool live = f0 + fr < a.n_frames;
      const float re = x[q].x, im = x[q].y;
      const float p2 = re * re + im * im;
      const float mg = sqrtf(p2);
      // d loss / d (magnitude^power)
      float gm = 0.f;
      if (a.n_mels > 0) {
        const int e0 = a.fbt_ptr[k], e1 = a.fbt_ptr[k + 1];
        for (int e = e0; e < e1; ++e) gm = fmaf(a.fbt_w[e], sh.melS[a.fbt_mel[e] * FR + fr], gm);
      }
      if (d_mag && live) gm += d_mag[((int64_t)b * K + k) * a.n_frames + f0 + fr];
      float gre = 0.f, gim = 0.f;
      if (a.power == 2) {
        gre = 2.f * gm * re;
        gim = 2.f * gm * im;
      } else if (mg > 0.f) {
        gre = gm * re / mg;
        gim = gm * im / mg;
      }
      if (d_phase && live && mg > a.phase_floor) {
        const float gp = d_phase[((int64_t)b * K + k) * a.n_frames + f0 + fr] / p2;
        gre = fmaf(-gp, im, gre);
        gim = fmaf(gp, re, gim);
      }
      if (!live) gre = gim = 0.f;
      G[q] = make_float2(gre, gim);
    }
    if (k == 0 || k == N / 2) {
      // H = (G1.re) + i (G2.re); conj(H) = (G1.re, -G2.re)
      z[k] = make_float2(G[0].x, -G[1].x);
    } else {
      // H[k] = G1/2 + i G2/2 ; H[N-k] = conj(G1)/2 + i conj(G2)/2
      const float2 hk = make_float2(0.5f * (G[0].x - G[1].y), 0.5f * (G[0].y + G[1].x));
      const float2 hn = make_float2(0.5f * (G[0].x + G[1].y), 0.5f * (-G[0].y + G[1].x));
      z[k] = make_float2(hk.x, -hk.y);
      z[N - k] = make_float2(hn.x, -hn.y);
    }
  }
  __syncthreads();
  const int grp = tid / GROUP, gt = tid - grp * GROUP;
  fft_dif(sh.z + grp * N, sh.tw, logn, gt);
  // y = FFT(conj H) (bit-reversed); dx1[n] = y.x, dx2[n] = -y.y; window, reflect, scatter
  float* __restrict__ db = d_audio + (int64_t)b * d_audio_bs;
  for (int idx = tid; idx < PAIRS * N; idx += blockDim.x) {
    const int p = idx / N, n = idx - p * N;
    const int fa = f0 + 2 * p;
    if (fa >= a.n_frames) continue;
    const int r = (int)(__brev((unsigned)n) >> (32 - logn));
    const float2 y = sh.z[p * N + r];
    const float w = a.window[n];
    if (w == 0.f) continue;
    atomicAdd(db + reflect_index(fa * a.hop + n - N / 2, a.L), w * y.x);
    if (fa + 1 < a.n_frames) atomicAdd(db + reflect_index((fa + 1) * a.hop + n - N / 2, a.L), -w * y.y);
  }
}

int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return (1 << l) == n ? l : -1;
}

int check_spec_args(const sty_spectrogram_args* a, const char* who) {
  STY_REQUIRE(a && a->audio && a->window && a->twiddle, "%s: null pointer", who);
  const int logn = ilog2_exact(a->n_fft);
  STY_REQUIRE(logn >= 8 && logn <= 12, "%s: n_fft must be a power of two in [256,4096], got %d", who, a->n_fft);
  STY_REQUIRE(a->B > 0 && a->L > a->n_fft / 2 && a->hop > 0, "%s: bad B/L/hop (%d,%d,%d)", who, a->B, a->L, a->hop);
  STY_REQUIRE(a->n_frames > 0 && a->n_frames <= a->L / a->hop + 1, "%s: n_frames %d out of range", who, a->n_frames);
  STY_REQUIRE(a->power == 1 || a->power == 2, "%s: power must be 1 or 2", who);
  STY_REQUIRE(a->n_mels == 0 || (a->fb_start && a->fb_len && a->fb_off && a->fb_w), "%s: mel filterbank missing", who);
  STY_REQUIRE(a->mel_mode >= 0 && a->mel_mode <= 2, "%s: bad mel_mode", who);
  return STY_OK;
}

}  // namespace

}  // namespace sty

using namespace sty;

extern "C" int sty_spectrogram_fwd(const sty_spectrogram_args* a, sty_stream_t stream) {
  int rc = check_spec_args(a, "sty_spectrogram_fwd");
  if (rc) return rc;
  STY_REQUIRE(!a->mel || a->n_mels > 0, "sty_spectrogram_fwd: mel output without a filterbank");
  constexpr int PAIRS = 4;
  const int logn = ilog2_exact(a->n_fft);
  const size_t smem = spec_smem_bytes<PAIRS>(a->n_fft, a->n_mels);
  STY_REQUIRE(smem <= 220 * 1024, "sty_spectrogram_fwd: shared memory %zu too large", smem);
  cudaFuncSetAttribute(spectrogram_fwd_kernel<PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(a->n_frames, 2 * PAIRS), a->B);
  spectrogram_fwd_kernel<PAIRS><<<grid, PAIRS * GROUP, smem, as_stream(stream)>>>(*a, logn);
  STY_CHECK_LAUNCH("sty_spectrogram_fwd");
  return STY_OK;
}

extern "C" int sty_spectrogram_bwd(const sty_spectrogram_args* a, const float* d_mel, const float* d_phase,
                                   const float* d_mag, float* d_audio, int64_t d_audio_bs,
                                   sty_stream_t stream) {
  int rc = check_spec_args(a, "sty_spectrogram_bwd");
  if (rc) return rc;
  STY_REQUIRE(d_audio, "sty_spectrogram_bwd: d_audio is null");
  STY_REQUIRE(a->n_mels == 0 || (a->fbt_ptr && a->fbt_mel && a->fbt_w), "sty_spectrogram_bwd: transposed filterbank missing");
  constexpr int PAIRS = 4;
  const int logn = ilog2_exact(a->n_fft);
  const size_t smem = spec_smem_bytes<PAIRS>(a->n_fft, a->n_mels);
  STY_REQUIRE(smem <= 220 * 1024, "sty_spectrogram_bwd: shared memory %zu too large", smem);
  cudaFuncSetAttribute(spectrogram_bwd_kernel<PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(cdiv(a->n_frames, 2 * PAIRS), a->B);
  spectrogram_bwd_kernel<PAIRS><<<grid, PAIRS * GROUP, smem, as_stream(stream)>>>(*a, d_mel, d_phase, d_mag, d_audio,
                                                                                   d_audio_bs, logn);
  STY_CHECK_LAUNCH("sty_spectrogram_bwd");
  return STY_OK;
}

// ---------------------------------------------------------------------------
// log-energy of a normalised log-mel:  log( || exp(mel*std+mean) ||_2 over mel bins + 1e-9 )
namespace sty {
namespace {
__global__ void __launch_bounds__(256)
mel_energy_kernel(const float* __restrict__ mel, float* __restrict__ out, int n_mels, int F, float mean,
                  float std, float eps) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float* __restrict__ mb = mel + (int64_t)b * n_mels * F;
  float acc = 0.f;
  for (int m = 0; m < n_mels; ++m) {
    const float v = expf(fmaf(mb[(int64_t)m * F + f], std, mean));
    acc = fmaf(v, v, acc);
  }
  out[(int64_t)b * F + f] = logf(sqrtf(acc) + eps);
}

// ---- L1 spectral convergence: sums[0] += sum|t-p|, sums[1] += sum|t|
__global__ void __launch_bounds__(256)
l1_sums_kernel(const float* __restrict__ t, const float* __restrict__ p, int64_t n, float* __restrict__ sums) {
  __shared__ float red[32];
  float a = 0.f, c = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float tv = t[i];
    a += fabsf(tv - p[i]);
    c += fabsf(tv);
  }
  a = block_sum(a, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    atomicAdd(sums, a);
    atomicAdd(sums + 1, c);
  }
}

// d_p[i] = coef[0] * sign(p - t)      (coef = g / (sum|t| + 1e-6), on the device)
__global__ void __launch_bounds__(256)
l1_bwd_kernel(const float* __restrict__ t, const float* __restrict__ p, int64_t n, const float* __restrict__ coef,
              float* __restrict__ d_p) {
  const float c = coef[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = p[i] - t[i];
    d_p[i] = d > 0.f ? c : (d < 0.f ? -c : 0.f);
  }
}

__device__ __forceinline__ float wrap_pi(float d) { return d - 6.283185307179586f * rintf(d * 0.15915494309189535f); }
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// anti-wrapping phase loss sums over (B,K,N): sums[0] = sum w_k aw(d), sums[1] = sum_{k<K-1} w_k aw(d[k+1]-d[k]),
// sums[2] = sum_{n<N-1} w_k aw(d[n+1]-d[n]),  d = pred - target,  w_k = base^k
__global__ void __launch_bounds__(256)
phase_sums_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int K, int N, float log_base,
                  float* __restrict__ sums) {
  __shared__ float red[32];
  const int bk = blockIdx.x;  // b*K + k
  const int k = bk % K;
  const float w = expf(log_base * (float)k);
  const float* __restrict__ p0 = pred + (int64_t)bk * N;
  const float* __restrict__ t0 = tgt + (int64_t)bk * N;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float d = p0[n] - t0[n];
    s0 += fabsf(wrap_pi(d));
    if (k + 1 < K) s1 += fabsf(wrap_pi((p0[N + n] - t0[N + n]) - d));
    if (n + 1 < N) s2 += fabsf(wrap_pi((p0[n + 1] - t0[n + 1]) - d));
  }
  s0 = block_sum(s0 * w, red);
  s1 = block_sum(s1 * w, red);
  s2 = block_sum(s2 * w, red);
  if (threadIdx.x == 0) {
    atomicAdd(sums, s0);
    atomicAdd(sums + 1, s1);
    atomicAdd(sums + 2, s2);
  }
}

// d_pred[b,k,n] of  coef[0]*mean-form sums: c0 = g/(B K N), c1 = g/(B (K-1) N), c2 = g/(B K (N-1)) with
// g = coef[0] read from the device
__global__ void __launch_bounds__(256)
phase_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int B, int K, int N, float log_base,
                 const float* __restrict__ coef, float* __restrict__ d_pred) {
  const int bk = blockIdx.x;
  const int k = bk % K;
  const float g = coef[0];
  const float c0 = g / ((float)B * K * N), c1 = g / ((float)B * (K - 1) * N), c2 = g / ((float)B * K * (N - 1));
  const float w = expf(log_base * (float)k);
  const float wm = expf(log_base * (float)(k - 1));
  const float* __restrict__ p0 = pred + (int64_t)bk * N;
  const float* __restrict__ t0 = tgt + (int64_t)bk * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float d = p0[n] - t0[n];
    float acc = c0 * w * sgn(wrap_pi(d));
    if (k + 1 < K) acc -= c1 * w * sgn(wrap_pi((p0[N + n] - t0[N + n]) - d));
    if (k > 0) acc += c1 * wm * sgn(wrap_pi(d - (p0[n - N] - t0[n - N])));
    if (n + 1 < N) acc -= c2 * w * sgn(wrap_pi((p0[n + 1] - t0[n + 1]) - d));
    if (n > 0) acc += c2 * w * sgn(wrap_pi(d - (p0[n - 1] - t0[n - 1])));
    d_pred[(int64_t)bk * N + n] = acc;
  }
}

// loss bookkeeping on the device (no host sync): see header
__global__ void stft_loss_finalize_kernel(const float* __restrict__ l1_sums, const float* __restrict__ ph_sums,
                                          const float* __restrict__ ph_counts, int n_res, float w_mel, float w_phase,
                                          int normalize, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mel = 0.f, ph = 0.f;
  for (int r = 0; r < n_res; ++r) {
    mel += l1_sums[2 * r] / (l1_sums[2 * r + 1] + 1e-6f);
    ph += ph_sums[3 * r] / ph_counts[3 * r] + ph_sums[3 * r + 1] / ph_counts[3 * r + 1] +
          ph_sums[3 * r + 2] / ph_counts[3 * r + 2];
  }
  mel /= (float)n_res;
  ph /= (float)n_res;
  const float gm = normalize ? w_mel / (mel + 1e-9f) : w_mel;
  const float gp = normalize ? w_phase / (ph + 1e-9f) : w_phase;
  out[0] = mel;
  out[1] = ph;
  out[2] = gm * mel + gp * ph;
  for (int r = 0; r < n_res; ++r) {
    out[4 + r] = gm / (float)n_res / (l1_sums[2 * r + 1] + 1e-6f);
    out[4 + n_res + r] = gp / (float)n_res;
  }
}
}  // namespace
}  // namespace sty

extern "C" int sty_mel_energy_fwd(const float* mel, float* out, int B, int n_mels, int F, float mean, float std,
                                  sty_stream_t stream) {
  STY_REQUIRE(mel && out && B > 0 && n_mels > 0 && F > 0, "sty_mel_energy_fwd: bad arguments");
  dim3 grid(cdiv(F, 256), B);
  mel_energy_kernel<<<grid, 256, 0, as_stream(stream)>>>(mel, out, n_mels, F, mean, std, 1e-9f);
  STY_CHECK_LAUNCH("sty_mel_energy_fwd");
  return STY_OK;
}

extern "C" int sty_l1_sums_fwd(const float* target, const float* pred, int64_t n, float* sums, sty_stream_t stream) {
  STY_REQUIRE(target && pred && sums && n > 0, "sty_l1_sums_fwd: bad arguments");
  const int grid = (int)(n / 1024 < 1 ? 1 : (n / 1024 > 1184 ? 1184 : n / 1024));
  l1_sums_kernel<<<grid, 256, 0, as_stream(stream)>>>(target, pred, n, sums);
  STY_CHECK_LAUNCH("sty_l1_sums_fwd");
  return STY_OK;
}

extern "C" int sty_l1_sums_bwd(const float* target, const float* pred, int64_t n, const float* coef, float* d_pred,
                               sty_stream_t stream) {
  STY_REQUIRE(target && pred && coef && d_pred && n > 0, "sty_l1_sums_bwd: bad arguments");
  const int grid = (int)(n / 1024 < 1 ? 1 : (n / 1024 > 1184 ? 1184 : n / 1024));
  l1_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(target, pred, n, coef, d_pred);
  STY_CHECK_LAUNCH("sty_l1_sums_bwd");
  return STY_OK;
}

extern "C" int sty_phase_loss_fwd(const float* pred, const float* target, int B, int K, int N, float* sums,
                                  sty_stream_t stream) {
  STY_REQUIRE(pred && target && sums && B > 0 && K > 1 && N > 1, "sty_phase_loss_fwd: bad arguments");
  const float log_base = logf(2.5f) / (float)(K / 2);
  phase_sums_kernel<<<B * K, 256, 0, as_stream(stream)>>>(pred, target, K, N, log_base, sums);
  STY_CHECK_LAUNCH("sty_phase_loss_fwd");
  return STY_OK;
}

extern "C" int sty_phase_loss_bwd(const float* pred, const float* target, int B, int K, int N, const float* coef,
                                  float* d_pred, sty_stream_t stream) {
  STY_REQUIRE(pred && target && coef && d_pred && B > 0 && K > 1 && N > 1, "sty_phase_loss_bwd: bad arguments");
  const float log_base = logf(2.5f) / (float)(K / 2);
  phase_bwd_kernel<<<B * K, 256, 0, as_stream(stream)>>>(pred, target, B, K, N, log_base, coef, d_pred);
  STY_CHECK_LAUNCH("sty_phase_loss_bwd");
  return STY_OK;
}

extern "C" int sty_stft_loss_finalize(const float* l1_sums, const float* phase_sums, const float* phase_counts,
                                      int n_res, float w_mel, float w_phase, int normalize, float* out,
                                      sty_stream_t stream) {
  STY_REQUIRE(l1_sums && phase_sums && phase_counts && out && n_res > 0 && n_res <= 8,
              "sty_stft_loss_finalize: bad arguments");
  stft_loss_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(l1_sums, phase_sums, phase_counts, n_res, w_mel, w_phase,
                                                             normalize, out);
  STY_CHECK_LAUNCH("sty_stft_loss_finalize");
  return STY_OK;
}
