// TMA (cp.async.bulk.tensor) helpers: device-side PTX wrappers and the host-side tensor-map factory.
// The driver entry point cuTensorMapEncodeTiled is fetched at run time with cudaGetDriverEntryPoint, so the
// library has no link-time dependency on libcuda (it must dlopen on a box without a driver for symbol checks).
#pragma once
#include <cuda.h>

#include "umma.cuh"

namespace sty {

// ------------------------------------------------------------------ device side
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global (tensor map, coordinates fastest-first) -> shared, completion on an mbarrier (complete_tx::bytes).
// Out-of-bounds elements of the box (negative or >= dim coordinates) are written as zeros.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global (tensor map); elements of the box outside the tensor are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ host side
// fp32 (B, C, T) activation with element strides (bs, cs, 1) as a 3-D tensor map (T fastest) with box
// (box_t, box_c, 1), no swizzle, zero fill.  Requirements of the hardware: base 16-byte aligned, byte strides
// multiples of 16, box_t * 4 a multiple of 16, box dims <= 256.  Returns false when they do not hold (callers
// then use the load/store path) or when the driver entry point is unavailable.
bool make_tmap_bct(CUtensorMap* out, const float* base, int64_t B, int64_t C, int64_t T, int64_t bs, int64_t cs,
                   int box_t, int box_c);
bool tma_layout_ok(const float* base, int64_t bs, int64_t cs);

}  // namespace sty
