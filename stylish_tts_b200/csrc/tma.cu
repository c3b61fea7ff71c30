// Host side of tma.cuh: tensor maps for (B, C, T) fp32 activations, cached per (pointer, shape, box).
#include <mutex>
#include <unordered_map>

#include "tma.cuh"

namespace sty {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool tma_layout_ok(const float* base, int64_t bs, int64_t cs) {
  return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (bs & 3) == 0 && (cs & 3) == 0 && cs > 0 && bs >= 0;
}

namespace {
struct Key {
  const void* base;
  int64_t B, C, T, bs, cs;
  int box_t, box_c;
  bool operator==(const Key& o) const {
    return base == o.base && B == o.B && C == o.C && T == o.T && bs == o.bs && cs == o.cs && box_t == o.box_t &&
           box_c == o.box_c;
  }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (int64_t v : {k.B, k.C, k.T, k.bs, k.cs, (int64_t)k.box_t, (int64_t)k.box_c})
      h = (h ^ (uint64_t)v) * 0x100000001B3ull;
    return (size_t)h;
  }
};
}  // namespace

bool make_tmap_bct(CUtensorMap* out, const float* base, int64_t B, int64_t C, int64_t T, int64_t bs, int64_t cs,
                   int box_t, int box_c) {
  if (!tma_layout_ok(base, bs, cs) || (box_t & 3) != 0 || box_t > 256 || box_c > 256 || box_t < 4 || box_c < 1)
    return false;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  // the encoding is a pure function of its arguments: cache it (the caching allocator of the host framework
  // hands back the same pointers every step, so an eager loop pays the ~1 us driver call once per site)
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  const Key key{base, B, C, T, bs, cs, box_t, box_c};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return true;
    }
  }
  // a batch of one (or a zero batch stride) still needs a legal outer stride
  const int64_t bs_eff = (B > 1 && bs > 0) ? bs : cs * C;
  cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)cs * 4, (cuuint64_t)bs_eff * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_t, (cuuint32_t)box_c, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return true;
}

}  // namespace sty
