// Shared device/host helpers for the cfm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "../../include/cfm_b200.h"

namespace cfm {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define CFM_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::cfm::set_error(__VA_ARGS__);             \
      return -1;                                 \
    }                                            \
  } while (0)

#define CFM_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::cfm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                       __FILE__, __LINE__);                                          \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

// every kernel launch goes through this so that bench.py can report `gpu_launches`; the per-kernel counters
// (cfm_kernel_launches("ffn_fused") ...) let the tests assert WHICH engine served a call
void count_launch(const char* name);
void count_variant(const char* name);
#define CFM_LAUNCHED_K(name)                            \
  do {                                                  \
    ::cfm::count_launch(name);                          \
    CFM_CUDA_OK(cudaPeekAtLastError());                 \
  } while (0)

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct Act;
template <> struct Act<float> {
  static constexpr int kId = CFM_F32;
  static constexpr int kVec = 4;  // elements per 16 bytes
};
template <> struct Act<__nv_bfloat16> {
  static constexpr int kId = CFM_BF16;
  static constexpr int kVec = 8;
};

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// exact-ish activations for the fp32 path (match ATen's expf-based SiLU / sigmoid to ~1 ulp)
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// fast variants for the bf16 path: one MUFU (tanh.approx) per element; rel. error ~2^-11,
// far below bf16's 2^-9 output rounding.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu_fast(float x) {
  float h = 0.5f * x;
  return fmaf(h, tanh_fast(h), h);
}

template <typename T> __device__ __forceinline__ float act_silu(float x);
template <> __device__ __forceinline__ float act_silu<float>(float x) { return silu_f(x); }
template <> __device__ __forceinline__ float act_silu<__nv_bfloat16>(float x) { return silu_fast(x); }
template <typename T> __device__ __forceinline__ float act_sigmoid(float x);
template <> __device__ __forceinline__ float act_sigmoid<float>(float x) { return sigmoid_f(x); }
template <> __device__ __forceinline__ float act_sigmoid<__nv_bfloat16>(float x) { return sigmoid_fast(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the layer stack is launched with programmaticStreamSerialization: its CTAs may be scheduled while the
// previous kernel drains, run their prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in
// griddepcontrol.wait until the previous grid has completed and flushed.  Hides launch latency + prologue at each of
// the ~100 kernel boundaries of a forward pass (works inside CUDA-graph capture: programmatic dependency edges).
// RULE: a kernel launched through launch_pdl() must execute pdl_wait() before its first global-memory access.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// environment switches are read once per process (thread-safe function-local statics)
inline bool env_is(const char* name, const char* value) {
  const char* e = getenv(name);
  return e != nullptr && strcmp(e, value) == 0;
}
inline bool pdl_enabled() {
  static const bool v = !env_is("CFM_B200_PDL", "0");
  return v;
}

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
// cudaFuncSetAttribute is a PER-DEVICE setting: remember which devices of this process already have it
#define CFM_SMEM_OPT_IN(kernel, bytes)                                                                        \
  do {                                                                                                        \
    static std::atomic<unsigned long long> _done{0};                                                          \
    const int _dev = ::cfm::current_device() & 63;                                                            \
    if (!((_done.load(std::memory_order_acquire) >> _dev) & 1ull)) {                                          \
      CFM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      _done.fetch_or(1ull << _dev, std::memory_order_release);                                                \
    }                                                                                                         \
  } while (0)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int num_sms() {
  static std::atomic<int> cache[64];
  const int dev = current_device() & 63;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---------------------------------------------------------------- engine entry points (defined per .cu)
int gemm_simt(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N,
              int K, int dtype, int epilogue, const float* residual, float alpha,
              const uint8_t* row_valid, cudaStream_t st);
// returns 1 when the shape/dtype is handled by the tcgen05 kernel
bool gemm_tc_supported(int lda, int ldc, int M, int N, int K, int dtype, int epilogue);
int gemm_tc(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N,
            int K, int dtype, int epilogue, const float* residual, float alpha,
            const uint8_t* row_valid, cudaStream_t st);
int gemm_tc_argmax(const void* A, int lda, const void* W, const float* bias, int M, int V, int K, unsigned long long* keys,
                   cudaStream_t st);
int gemm_tc_init();
// RESIDUAL epilogue with fused LayerNorm(s) (ln_mode 1 or 2); needs N == 256 (one tile per row)
bool gemm_tc_ln_supported(int lda, int ldx, int ldy, int M, int N, int K, int dtype);
int gemm_tc_ln(const void* A, int lda, const void* W, const float* bias, float* X, int ldx, const float* residual, int M,
               int N, int K, float alpha, const uint8_t* row_valid, int ln_mode, const float* g1, const float* b1,
               const float* g2, const float* b2, void* Y, int ldy, const uint8_t* y_row_valid, float eps, int epilogue,
               void* Cact, int ldc, cudaStream_t st);

bool ffn_fused_supported(int ld_in, int ldx, int ld_out, int M, int d, int F, int dtype, int ln_mode);
int ffn_fused(const void* y_in, int ld_in, const void* W1, const float* b1, const void* W2, const float* b2, float* X,
              int ldx, int M, int F, float alpha, int ln_mode, const float* g1, const float* be1, const float* g2,
              const float* be2, void* y_out, int ld_out, const uint8_t* y_row_valid, float eps, cudaStream_t st);

struct FfnModule {               // one PositionwiseFeedForwardModule + the LayerNorm(s) after its residual add
  const void* W1; const float* b1; const void* W2; const float* b2;
  float alpha;
  const float* g1; const float* be1; const float* g2; const float* be2;
};
bool ffn_chain_supported(int M, int d, int F, int dtype, const FfnModule* a, const FfnModule& b, int Np);
int ffn_chain(const void* y_in, const FfnModule* a, const FfnModule& b, float* X, int M, int F, void* y_out,
              const uint8_t* y_row_valid, float eps, const void* Wp, const float* bp, void* P, int Np, cudaStream_t st);
bool mhsa_fused_supported(int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts, int B, int H,
                          int Tq, int Tk, int d, int dtype, bool has_key_bias);
int mhsa_fused(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts, const void* v,
               int64_t v_bs, int64_t v_ts, int B, int T, const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, float scale,
               const void* Wo, const float* bo, float* X, const float* g1, const float* be1, void* Y,
               const uint8_t* y_row_valid, float eps, cudaStream_t st);
bool conv_fused_supported(int M, int T, int d, int k, int dtype);
int conv_fused(const void* y_in, const void* W1, const float* b1, const float* dw_w, const float* dw_b, const void* W2,
               const float* b2, float* X, int M, int T, const uint8_t* row_valid, const float* g1, const float* be1,
               void* y_out, float eps, cudaStream_t st);
int attention_simt(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts,
                   const void* v, int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk,
                   const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, const float* key_bias,
                   float scale, int dtype, cudaStream_t st);
bool attention_tc_supported(int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs,
                            int64_t v_ts, int B, int H, int Tq, int Tk, int dtype);
int attention_tc(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts,
                 const void* v, int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk,
                 const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, const float* key_bias,
                 float scale, int dtype, cudaStream_t st);
int attention_tc_init();
// two-query-tile ping-pong organisation for long query ranges (attention_pp.cu); same contract as attention_tc
int attention_pp(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts, const void* v,
                 int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk, const uint8_t* mask, int64_t mask_bs,
                 int64_t mask_rs, float scale, cudaStream_t st);

// general GEMM of the training path (gemm_gen.cu): C (+)= alpha * opA(A) opB(B)^T, transposed operands, (head, batch) dims
bool gemm_gen_tc_supported(const void* A, long long lda, long long a_hs, long long a_bs, const void* B, long long ldb,
                           long long b_hs, long long b_bs, const void* C, int c_dtype, long long ldc, long long c_hs,
                           long long c_bs, int M, int N, int K, int in_dtype);
int gemm_gen(const void* A, int a_mn, long long lda, long long a_hs, long long a_bs, const void* B, int b_mn, long long ldb,
             long long b_hs, long long b_bs, void* C, int c_dtype, long long ldc, long long c_hs, long long c_bs,
             int accumulate, int M, int N, int K, int nH, int nB, float alpha, int splits, cudaStream_t st);
int gemm_gen_simt(const void* A, int a_mn, long long lda, long long a_hs, long long a_bs, const void* B, int b_mn,
                  long long ldb, long long b_hs, long long b_bs, void* C, int c_dtype, long long ldc, long long c_hs,
                  long long c_bs, int accumulate, int M, int N, int K, int nH, int nB, int in_dtype, float alpha,
                  cudaStream_t st);

}  // namespace cfm
