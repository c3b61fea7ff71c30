// standalone TMA sanity test: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_test tma_test.cu ../../stylish_tts_b200/csrc/tma.cu ../../stylish_tts_b200/csrc/err.cu
#include <vector>
#include "../../stylish_tts_b200/csrc/tma.cuh"
using namespace sty;

__global__ void k_param(const __grid_constant__ CUtensorMap tmap, float* out, int box_t, int box_c, int t0, int b) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* dst = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + box_t * box_c * 4);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, box_t * box_c * 4);
    tma_load_3d(dst, &tmap, bar, t0, 0, b);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < box_t * box_c; i += blockDim.x) out[i] = dst[i];
}
__global__ void k_ptr(const CUtensorMap* tmap, float* out, int box_t, int box_c, int t0, int b) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* dst = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + box_t * box_c * 4);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, box_t * box_c * 4);
    tma_load_3d(dst, tmap, bar, t0, 0, b);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < box_t * box_c; i += blockDim.x) out[i] = dst[i];
}

int run(int B, int C, int T, int pitch, int box_t, int box_c, int t0, int b, bool by_ptr) {
  std::vector<float> h((size_t)B * C * pitch);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, (size_t)box_t * box_c * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  if (!make_tmap_bct(&m, d, B, C, T, (int64_t)C * pitch, pitch, box_t, box_c)) { printf("encode failed\n"); return 1; }
  size_t smem = (size_t)box_t * box_c * 4 + 64;
  if (by_ptr) {
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_ptr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_ptr<<<1, 128, smem>>>(dm, o, box_t, box_c, t0, b);
  } else {
    cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_param<<<1, 128, smem>>>(m, o, box_t, box_c, t0, b);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> r((size_t)box_t * box_c);
  cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < box_c; ++c)
    for (int i = 0; i < box_t; ++i) {
      int t = t0 + i;
      float want = (t >= 0 && t < T && c < C) ? h[((size_t)b * C + c) * pitch + t] : 0.f;
      if (r[c * box_t + i] != want) ++bad;
    }
  printf("  B=%d C=%d T=%d pitch=%d box=(%d,%d) t0=%d b=%d by_ptr=%d -> %d mismatches\n", B, C, T, pitch, box_t, box_c, t0, b, (int)by_ptr, bad);
  return bad ? 3 : 0;
}
int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0, by_ptr = argc > 2 ? atoi(argv[2]) : 0;
  switch (mode) {
    case 0: return run(2, 32, 1000, 1000, 148, 32, -10, 0, by_ptr);
    case 1: return run(2, 32, 1000, 1000, 128, 32, 0, 1, by_ptr);
    case 2: return run(2, 32, 60225, 60228, 140, 32, 60160 - 5, 1, by_ptr);
    case 3: return run(2, 64, 1000, 1000, 64, 64, 100, 1, by_ptr);
    case 4: return run(2, 32, 1000, 1000, 32, 32, 0, 1, by_ptr);
    case 5: return run(2, 32, 1024, 1024, 64, 16, 64, 1, by_ptr);
    case 6: return run(2, 32, 1000, 1000, 148, 32, 100, 1, by_ptr);
    case 7: return run(2, 32, 1000, 1000, 128, 32, -10, 1, by_ptr);
    case 8: return run(2, 32, 1000, 1000, 128, 32, 950, 1, by_ptr);
    case 9: return run(2, 32, 60225, 60228, 128, 32, 1280, 1, by_ptr);
    case 10: return run(2, 32, 60228, 60228, 128, 32, 60200, 1, by_ptr);
    case 11: return run(2, 32, 1000, 1000, 144, 32, 100, 1, by_ptr);
    case 12: return run(2, 32, 1000, 1000, 160, 32, 100, 1, by_ptr);
    default: return run(2, 32, 60225, 60228, mode, 32, 60160 - 5, 1, by_ptr);
  }
  return 0;
}
