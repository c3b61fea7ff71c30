// Dropout / Dropout1d / DropPath as one elementwise kernel with an optional fused activation, scale and
// residual; the keep mask is a stateless hash (common.cuh drop_keep), regenerated in the backward.
// Replaces nn.Dropout (text_encoder.py:63,323,352; conformer.py:90-92,144,187,245; ada_norm.py:157),
// nn.Dropout1d (duration_predictor.py:30,79) and DropPath (conv_next.py:138-153) in training mode.
#include <math.h>

#include "common.cuh"

namespace sty {
namespace {

template <bool BWD>
__global__ void __launch_bounds__(256)
dropout_kernel(const float* x, const float* aux, float* y, int64_t n,  // may alias (in-place use)
               int64_t group, int act, float scale, DropSpec ds) {
  const unsigned long long seed = ds.seed ? *ds.seed : 0ull;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t mi = group == 1 ? i : i / group;
    const float keep = (!ds.seed || drop_keep(seed, ds.site, (unsigned long long)mi, ds.thresh)) ? ds.inv_keep : 0.f;
    const float xv = x[i];
    if constexpr (!BWD) {
      const float a = act == STY_ACT_SWISH ? xv / (1.0f + expf(-xv)) : xv;
      y[i] = (aux ? aux[i] : 0.f) + scale * a * keep;
    } else {
      float d = 1.f;
      if (act == STY_ACT_SWISH) {
        const float sg = 1.0f / (1.0f + expf(-xv));
        d = sg * (1.0f + xv * (1.0f - sg));
      }
      y[i] = aux[i] * scale * d * keep;
    }
  }
}

// y[r,t] = act(scale[r] * x[r,t] + shift[r]) * keep(r*T + t) / (1-p): the AdaIN affine + LeakyReLU + Dropout in
// front of the convs of AdaptiveDecoderBlock in train() mode (ada_norm.py:181-186), materialised once.
__global__ void __launch_bounds__(256)
affine_act_dropout_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                          const float* __restrict__ shift, float* __restrict__ y, int T, int act, DropSpec ds) {
  const unsigned long long seed = ds.seed ? *ds.seed : 0ull;
  const int64_t r = blockIdx.y;
  const float sc = scale[r], sh = shift[r];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const int64_t i = r * T + t;
    const float keep = (!ds.seed || drop_keep(seed, ds.site, (unsigned long long)i, ds.thresh)) ? ds.inv_keep : 0.f;
    y[i] = act_apply(fmaf(sc, x[i], sh), act) * keep;
  }
}

}  // namespace
}  // namespace sty

using namespace sty;

extern "C" int sty_affine_act_dropout_fwd(const float* x, const float* scale, const float* shift, float* y,
                                          int rows, int T, int act, const sty_dropout* drop, sty_stream_t stream) {
  STY_REQUIRE(x && scale && shift && y && rows > 0 && T > 0 && rows <= 65535, "affine_act_dropout: bad argument");
  STY_REQUIRE(act == STY_ACT_NONE || act == STY_ACT_RELU || act == STY_ACT_LEAKY02 || act == STY_ACT_SWISH ||
                  act == STY_ACT_GELU, "affine_act_dropout: activation %d needs a parameter", act);
  STY_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "affine_act_dropout: p must be in [0,1)");
  dim3 grid(cdiv(T, 256) > 64 ? 64 : cdiv(T, 256), rows);
  affine_act_dropout_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, scale, shift, y, T, act, make_drop(drop));
  STY_CHECK_LAUNCH("affine_act_dropout");
  return STY_OK;
}

static int dropout_launch(bool bwd, const float* x, const float* aux, float* y, int64_t n, int64_t group, int act,
                          float scale, const sty_dropout* drop, sty_stream_t stream) {
  STY_REQUIRE(x && y && n > 0 && group >= 1, "dropout: bad argument");
  STY_REQUIRE(act == STY_ACT_NONE || act == STY_ACT_SWISH, "dropout: activation %d not built (none, swish)", act);
  STY_REQUIRE(!drop || (drop->p >= 0.f && drop->p < 1.f), "dropout: p must be in [0,1)");
  STY_REQUIRE(!bwd || aux, "dropout_bwd: null dy");
  const DropSpec ds = make_drop(drop);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (bwd)
    dropout_kernel<true><<<(int)blocks, 256, 0, as_stream(stream)>>>(x, aux, y, n, group, act, scale, ds);
  else
    dropout_kernel<false><<<(int)blocks, 256, 0, as_stream(stream)>>>(x, aux, y, n, group, act, scale, ds);
  STY_CHECK_LAUNCH("dropout");
  return STY_OK;
}

extern "C" int sty_dropout_fwd(const float* x, const float* res, float* y, int64_t n, int64_t group, int act,
                               float scale, const sty_dropout* drop, sty_stream_t stream) {
  return dropout_launch(false, x, res, y, n, group, act, scale, drop, stream);
}

extern "C" int sty_dropout_bwd(const float* x, const float* dy, float* dx, int64_t n, int64_t group, int act,
                               float scale, const sty_dropout* drop, sty_stream_t stream) {
  return dropout_launch(true, x, dy, dx, n, group, act, scale, drop, stream);
}
