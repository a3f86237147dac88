// CTC head, greedy part (scope row f2): frame-wise argmax of ctc_lo(x) = x W^T + b over the vocabulary
// (decoder.py:14,19 of the reference defines the projection; greedy decoding = argmax, collapse repeats, drop blank 0).
// The (frames x vocab) logit matrix (317 MB in fp32 for the C2 batch) is never written on the tcgen05 engine: the GEMM's
// epilogue reduces every 128 x 256 tile to per-row (max, argmax) pairs and combines tiles with a 64-bit atomicMax
// (gemm_tc.cu, EPI_ARGMAX).  The CUDA-core engine (fp32 path) goes through a caller-provided logit workspace in chunks
// of rows.
#include "cfm_common.cuh"
#include <math_constants.h>
#include <algorithm>

namespace cfm {
namespace {

constexpr int kChunkRows = 1024;

__global__ void ctc_unpack_kernel(const unsigned long long* __restrict__ keys, int32_t* __restrict__ ids,
                                  float* __restrict__ best, int M) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const unsigned long long k = keys[i];
  uint32_t u = static_cast<uint32_t>(k >> 32);
  ids[i] = static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(k & 0xFFFFFFFFu));
  if (best != nullptr) {
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    best[i] = __uint_as_float(u);
  }
}

// one warp per row of a (rows, V) logit matrix: first index of the maximum
template <typename T>
__global__ void ctc_argmax_rows_kernel(const T* __restrict__ logits, int ld, int rows, int V, int32_t* __restrict__ ids,
                                       float* __restrict__ best) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* p = logits + (size_t)row * ld;
  float bv = -CUDART_INF_F;
  int bi = 0x7fffffff;
  for (int c = lane; c < V; c += 32) {
    const float x = to_f32(p[c]);
    if (x > bv) { bv = x; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) {
    ids[row] = bi == 0x7fffffff ? 0 : bi;
    if (best != nullptr) best[row] = bv;
  }
}

}  // namespace
}  // namespace cfm

extern "C" int64_t cfm_ctc_ws_bytes(int M, int V, int dtype) {
  const int64_t keys = (int64_t)M * 8;
  const int64_t logits = (int64_t)std::min(M, cfm::kChunkRows) * V * (dtype == CFM_BF16 ? 2 : 4);
  return std::max(keys, logits);
}

extern "C" int cfm_ctc_argmax(const void* x, int ldx, const void* W, const float* bias, int M, int V, int d, int dtype,
                              int32_t* ids, float* best, void* ws, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(x && W && ids && ws, "cfm_ctc_argmax: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_ctc_argmax: bad dtype %d", dtype);
  CFM_CHECK_ARG(M >= 0 && V > 0 && d > 0 && ldx >= d, "cfm_ctc_argmax: bad shape");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = gemm_tc_supported(ldx, 8, M, 256, d, dtype, CFM_EPI_BIAS);
  if (engine == CFM_ENGINE_TC) CFM_CHECK_ARG(tc_ok, "cfm_ctc_argmax: tcgen05 engine does not support M=%d d=%d dtype=%d", M, d, dtype);
  if (tc_ok && engine != CFM_ENGINE_SIMT) {
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
    CFM_CUDA_OK(cudaMemsetAsync(keys, 0, (size_t)M * 8, st));
    int rc = gemm_tc_argmax(x, ldx, W, bias, M, V, d, keys, st);
    if (rc != 0) return rc;
    ctc_unpack_kernel<<<(M + 255) / 256, 256, 0, st>>>(keys, ids, best, M);
    CFM_LAUNCHED_K("ctc_unpack");
    return 0;
  }
  const size_t esz = dtype == CFM_BF16 ? 2 : 4;
  for (int m0 = 0; m0 < M; m0 += kChunkRows) {
    const int rows = std::min(kChunkRows, M - m0);
    const void* xa = static_cast<const uint8_t*>(x) + (size_t)m0 * ldx * esz;
    int rc = cfm_gemm(xa, ldx, W, bias, ws, V, rows, V, d, dtype, CFM_EPI_BIAS, nullptr, 1.f, nullptr, CFM_ENGINE_SIMT, stream);
    if (rc != 0) return rc;
    if (dtype == CFM_BF16)
      ctc_argmax_rows_kernel<__nv_bfloat16><<<(rows + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)ws, V, rows, V, ids + m0,
                                                                            best ? best + m0 : nullptr);
    else
      ctc_argmax_rows_kernel<float><<<(rows + 7) / 8, 256, 0, st>>>((const float*)ws, V, rows, V, ids + m0,
                                                                    best ? best + m0 : nullptr);
    CFM_LAUNCHED_K("ctc_argmax_rows");
  }
  return 0;
}
