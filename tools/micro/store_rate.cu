// Micro-benchmark: how fast can one SM push a finished [128 x 256] fp32 tile + [128 x 256] bf16 tile (192 KB, the
// residual/LayerNorm epilogue's output) from shared memory to global memory?
//   mode 0: TMA bulk tensor stores of 128-byte-swizzled [128 rows x 128 B] boxes (what resid_epilogue.cuh does)
//   mode 1: coalesced st.global.v4 from the same shared memory (256 threads, one 512-byte row segment per warp)
//   mode 2: TMA loads of the fp32 tile only (128 KB) for comparison
// All CTAs run concurrently on distinct tiles (like the kernels).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_rate store_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../conformer_pytorch_lightning_b200/csrc/tc_common.cuh"
using namespace cfm::tc;
namespace cfm { namespace tc { EncodeTiledFn encode_tiled_fn() { return nullptr; } } }

__global__ void __launch_bounds__(256, 1) k(const __grid_constant__ CUtensorMap tmX, float* X, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < 12 * 16384 / 4; i += 256) ((uint32_t*)smem)[i] = i;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  const int m0 = blockIdx.x * 128;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      if (threadIdx.x == 0) {
        for (int c = 0; c < 12; ++c) { tma_store_2d(&tmX, smem + c * 16384, (c % 8) * 32, m0 + (c / 8) * 0); bulk_commit(); }
        bulk_wait_read<0>();
      }
      __syncthreads();
    } else if (mode == 1) {
      // 12 chunks of [128 rows x 128 B]: thread t writes 16 B; a warp covers 4 rows x 128 B (4 segments of 128 B)
      for (int c = 0; c < 12; ++c) {
        const uint8_t* src = smem + c * 16384;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int row = rr * 32 + (threadIdx.x >> 3), j = threadIdx.x & 7;
          uint4 v = *reinterpret_cast<const uint4*>(src + row * 128 + j * 16);
          *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(X) + ((size_t)(m0 + row) * 256 + (c % 8) * 32) * 4 + j * 16) = v;
        }
      }
      __syncthreads();
    } else {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, 8 * 16384);
        for (int c = 0; c < 8; ++c) tma_load_2d(smem + c * 16384, &tmX, &bar, c * 32, m0);
      }
      mbar_wait(&bar, it & 1);
    }
  }
  if (mode == 0 && threadIdx.x == 0) bulk_wait_all<0>();
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  const int M = 148 * 128;
  float* X; cudaMalloc(&X, (size_t)M * 256 * 4);
  long long* d; cudaMalloc(&d, 8);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  auto fn = reinterpret_cast<EncodeTiledFn>(fnp);
  CUtensorMap tm;
  cuuint64_t dims[2] = {256, (cuuint64_t)M}, str[1] = {256 * 4};
  cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384);
  const char* names[] = {"TMA store 192 KB", "st.global.v4 192 KB", "TMA load 128 KB"};
  for (int grid : {148, 1})
    for (int mode = 0; mode < 3; ++mode) {
      const int iters = 20;
      k<<<grid, 256, 12 * 16384>>>(tm, X, mode, iters, d);
      cudaDeviceSynchronize();
      k<<<grid, 256, 12 * 16384>>>(tm, X, mode, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double bytes = (mode == 2 ? 8 : 12) * 16384.0;
      printf("grid %3d  %-20s: %8.0f clk per tile  -> %5.1f B/clk/SM  [%s]\n", grid, names[mode], (double)h / iters,
             bytes * iters / h, cudaGetErrorString(e));
    }
  return 0;
}
