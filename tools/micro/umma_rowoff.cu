// Probe: can a SWIZZLE_128B K-major UMMA operand start at an arbitrary 128-byte row of a tile that was written with
// the absolute-address swizzle (chunk ^= row & 7)?  D[m][n] = A[m + off][n] through an identity B.  Tries the
// descriptor's base_offset field = 0 and = (start >> 7) & 7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rowoff umma_rowoff.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../conformer_pytorch_lightning_b200/csrc/tc_common.cuh"
using namespace cfm::tc;
namespace cfm { namespace tc { EncodeTiledFn encode_tiled_fn() { return nullptr; } } }

__device__ __forceinline__ float aval(int row, int k) { return (float)((row * 7 + k * 3) % 251 - 125); }

__global__ void __launch_bounds__(128, 1) k(int off, int use_base_off, int* mism) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* sA = smem;               // 256 rows x 128 B
  uint8_t* sB = smem + 32768;       // 64 rows x 128 B (identity)
  for (int i = threadIdx.x; i < 256 * 64; i += 128) {
    int row = i >> 6, kk = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(sA + row * 128 + ((((kk >> 3) ^ row) & 7) << 4) + (kk & 7) * 2) = __float2bfloat16(aval(row, kk));
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    int n = i >> 6, kk = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 128 + ((((kk >> 3) ^ n) & 7) << 4) + (kk & 7) * 2) = __float2bfloat16(n == kk ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<64>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm = slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = umma_idesc_bf16(128, 64);
    uint32_t a_addr = smem_u32(sA) + off * 128;
    uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(smem_u32(sB));
    if (use_base_off) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
    for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, da + 2 * kk, db + 2 * kk, idesc, kk != 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
  int bad = 0;
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) bad += (__uint_as_float(v[j]) != aval(m + off, c * 32 + j));
  }
  (void)lane;
  atomicAdd(mism, bad);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<64>(tm);
}

int main() {
  int* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  int offs[] = {0, 8, 1, 3, 4, 7, 20, 21, 41};
  for (int ub = 0; ub < 2; ++ub)
    for (int o : offs) {
      cudaMemset(d, 0, 4);
      k<<<1, 128, 49152>>>(o, ub, d);
      cudaError_t e = cudaDeviceSynchronize();
      int h = -1; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
      printf("row offset %2d base_offset_field=%d: %d mismatches of 8192 [%s]\n", o, ub, h, cudaGetErrorString(e));
    }
  return 0;
}
