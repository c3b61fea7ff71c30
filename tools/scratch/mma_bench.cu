// tcgen05.mma issue/throughput microbenchmark (K-major, SWIZZLE_NONE operands as conv1d_umma.cu stages them).
// nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o mma_bench mma_bench.cu
#include <cstdio>
#include <vector>
#include "../../stylish_tts_b200/csrc/umma.cuh"
using namespace sty;

// P: 0 N=64 same acc | 1 N=64 two accs alternating | 2 N=32 same acc | 3 pair (N=64, N=32) same acc
//    4 N=128 same acc | 5 N=256 same acc | 6 pair, accumulators alternate per pair | 7 N=16 same acc
//    8 N=96 same | 9 pair with a different A start address per MMA (tap shift) same acc
template <int P>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out, int rows) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 96 * 1024);
      const uint64_t a_d = make_desc(a_addr, (uint32_t)rows, 8u);
      const uint32_t a_hi = (uint32_t)(a_d >> 32), a0 = (uint32_t)a_d;
      constexpr int N = P == 2 ? 32 : P == 4 ? 128 : P == 5 ? 256 : P == 7 ? 16 : P == 8 ? 96 : 64;
      const uint32_t id = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t id32 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t b_d = make_desc(b_addr, (uint32_t)N, 8u);
      const uint32_t b_hi = (uint32_t)(b_d >> 32), b0 = (uint32_t)b_d;
      const uint32_t lo = 2 * rows;
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (P == 1) umma_bf16_w(tm + (u & 1) * 64, a0, a_hi, b0, b_hi, id, 1u);
          else if (P == 3) { umma_bf16_w(tm, a0, a_hi, b0, b_hi, id, 1u); umma_bf16_w(tm, a0 + lo, a_hi, b0, b_hi, id32, 1u); }
          else if (P == 6) { umma_bf16_w(tm + (u & 1) * 64, a0, a_hi, b0, b_hi, id, 1u); umma_bf16_w(tm + (u & 1) * 64, a0 + lo, a_hi, b0, b_hi, id32, 1u); }
          else if (P == 9) { umma_bf16_w(tm, a0 + u, a_hi, b0 + u * 128, b_hi, id, 1u); umma_bf16_w(tm, a0 + lo + u, a_hi, b0 + u * 128, b_hi, id32, 1u); }
          else umma_bf16_w(tm, a0, a_hi, b0, b_hi, id, 1u);
        }
      }
      t1 = clock64();
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  const long long t2 = clock64();
  if (tid == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int P>
void run(const char* name, long long* d, int grid) {
  const size_t smem = 160 * 1024;
  cudaFuncSetAttribute(bench<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4096;
  bench<P><<<grid, 128, smem>>>(16, d, 148);
  bench<P><<<grid, 128, smem>>>(iters, d, 148);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("grid %3d  %-22s issue %7.1f cyc/iter   complete %7.1f cyc/iter\n", grid, name, (double)h[0] / iters, (double)h[1] / iters);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  for (int grid : {1, 148}) {
    run<0>("N=64 same acc", d, grid);
    run<1>("N=64 two accs", d, grid);
    run<2>("N=32 same acc", d, grid);
    run<3>("pair(64,32) same acc", d, grid);
    run<4>("N=128 same acc", d, grid);
    run<5>("N=256 same acc", d, grid);
    run<6>("pair, accs alternate", d, grid);
    run<7>("N=16 same acc", d, grid);
    run<8>("N=96 same acc", d, grid);
    run<9>("pair, shifted A/B", d, grid);
  }
  return 0;
}
