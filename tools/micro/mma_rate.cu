// Micro-benchmark: sustained tcgen05.mma (kind::f16, cta_group::1, SS operands) issue rate per SM for N = 64/128/256,
// optionally with a concurrent TMA-like smem writer.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../conformer_pytorch_lightning_b200/csrc/tc_common.cuh"
using namespace cfm::tc;
namespace cfm { namespace tc { EncodeTiledFn encode_tiled_fn() { return nullptr; } } }

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  // zero operands
  for (int i = threadIdx.x; i < (16384 + 32768 * 2) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm = slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = umma_idesc_bf16(128, N);
    uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 16384));
    uint64_t db2 = umma_desc_sw128(smem_u32(smem + 16384 + 32768));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t b = (mode == 1 && (i & 1)) ? db2 : db;       // mode 1: alternate B buffers
        umma_bf16(tm + ((i & 1) ? N : 0) % 512, da + 2 * kk, b + 2 * kk, idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N> void run(const char* name, int mode, int grid) {
  long long* d; cudaMalloc(&d, 8);
  int iters = 2000;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  k<N><<<grid, 128, 100000>>>(iters, d, mode);
  cudaDeviceSynchronize();
  k<N><<<grid, 128, 100000>>>(iters, d, mode);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 4);
  printf("%s N=%d grid=%d: %.1f clk per MMA (ideal %d) -> %.0f MAC/clk/SM  [%s]\n", name, N, grid, per, N / 2, 128.0 * N * 16 / per,
         cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64>("ss", 0, 148); run<128>("ss", 0, 148); run<256>("ss", 0, 148);
  run<128>("ss-altB", 1, 148); run<256>("ss", 0, 1);
  return 0;
}
