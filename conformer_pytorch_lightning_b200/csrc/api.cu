// C-ABI surface of libcfm_b200.so: argument validation, engine dispatch, error reporting.
#include "cfm_common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

namespace cfm {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

// per-kernel launch counters: a small append-only table of (literal name, count)
struct KernelCounter { std::atomic<const char*> name{nullptr}; std::atomic<int64_t> n{0}; };
static KernelCounter g_kernels[96];

static void count_named(const char* name);
void count_launch(const char* name) {
  g_launches.fetch_add(1);
  count_named(name);
}
// a variant of a kernel family ("gemm_tc_pair" inside "gemm_tc"): named counter only, the launch itself is counted by
// count_launch under the family name
void count_variant(const char* name) { count_named(name); }
static void count_named(const char* name) {
  for (auto& k : g_kernels) {
    const char* cur = k.name.load(std::memory_order_acquire);
    if (cur == nullptr) {
      const char* expected = nullptr;
      if (k.name.compare_exchange_strong(expected, name, std::memory_order_acq_rel)) cur = name;
      else cur = expected;
    }
    if (cur == name || strcmp(cur, name) == 0) { k.n.fetch_add(1, std::memory_order_relaxed); return; }
  }
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace cfm

extern "C" int cfm_abi_version(void) { return CFM_ABI_VERSION; }
extern "C" const char* cfm_last_error(void) { return cfm::g_err; }
extern "C" int64_t cfm_launch_count(void) { return cfm::g_launches.load(); }
extern "C" int64_t cfm_kernel_launches(const char* name) {
  if (name == nullptr) return -1;
  int64_t total = 0;
  for (auto& k : cfm::g_kernels) {
    const char* cur = k.name.load(std::memory_order_acquire);
    if (cur == nullptr) break;
    if (strcmp(cur, name) == 0) total += k.n.load();
  }
  return total;
}

extern "C" int cfm_init(int device) {
  using namespace cfm;
  // per-device setup (shared-memory opt-ins are per device); the caller's current device is restored
  int prev = 0;
  CFM_CUDA_OK(cudaGetDevice(&prev));
  int major = 0, minor = 0;
  CFM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  CFM_CHECK_ARG(major == 10, "cfm_init: device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
  if (prev != device) CFM_CUDA_OK(cudaSetDevice(device));
  int rc = gemm_tc_init();
  if (rc == 0) rc = attention_tc_init();
  if (prev != device) cudaSetDevice(prev);
  return rc;
}

extern "C" int cfm_gemm(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N,
                        int K, int dtype, int epilogue, const float* residual, float alpha,
                        const uint8_t* row_valid, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(A && W && C, "cfm_gemm: null A/W/C");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_gemm: bad dtype %d", dtype);
  CFM_CHECK_ARG(epilogue >= CFM_EPI_BIAS && epilogue <= CFM_EPI_BIAS_RELU, "cfm_gemm: bad epilogue %d", epilogue);
  // EPI_RESIDUAL with residual == nullptr: fp32 output without a residual operand, X = alpha * rowmask(A W^T + b)
  CFM_CHECK_ARG(M >= 0 && N > 0 && K > 0 && lda >= K && ldc >= N, "cfm_gemm: bad shape M=%d N=%d K=%d lda=%d ldc=%d",
                M, N, K, lda, ldc);
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = gemm_tc_supported(lda, ldc, M, N, K, dtype, epilogue);
  if (engine == CFM_ENGINE_TC) {
    CFM_CHECK_ARG(tc_ok, "cfm_gemm: tcgen05 engine does not support M=%d N=%d K=%d dtype=%d epi=%d", M, N, K, dtype,
                  epilogue);
    return gemm_tc(A, lda, W, bias, C, ldc, M, N, K, dtype, epilogue, residual, alpha, row_valid, st);
  }
  if (engine == CFM_ENGINE_AUTO && tc_ok)
    return gemm_tc(A, lda, W, bias, C, ldc, M, N, K, dtype, epilogue, residual, alpha, row_valid, st);
  return gemm_simt(A, lda, W, bias, C, ldc, M, N, K, dtype, epilogue, residual, alpha, row_valid, st);
}

extern "C" int cfm_gemm_ln(const void* A, int lda, const void* W, const float* bias, float* X, int ldx, int M, int N,
                           int K, int dtype, float alpha, const uint8_t* row_valid, const float* g1, const float* b1,
                           const float* g2, const float* b2, void* Y, int ldy, const uint8_t* y_row_valid, float eps,
                           int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(A && W && X && Y && g1 && b1, "cfm_gemm_ln: null A/W/X/Y/g1/b1");
  CFM_CHECK_ARG((g2 == nullptr) == (b2 == nullptr), "cfm_gemm_ln: g2/b2 must both be set or both null");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_gemm_ln: bad dtype %d", dtype);
  CFM_CHECK_ARG(M >= 0 && N > 0 && K > 0 && lda >= K && ldx >= N && ldy >= N, "cfm_gemm_ln: bad shape");
  if (M == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused_ok = gemm_tc_ln_supported(lda, ldx, ldy, M, N, K, dtype);
  if (engine == CFM_ENGINE_TC)
    CFM_CHECK_ARG(fused_ok, "cfm_gemm_ln: fused tcgen05 path does not support M=%d N=%d K=%d dtype=%d", M, N, K, dtype);
  if (fused_ok && engine != CFM_ENGINE_SIMT)
    return gemm_tc_ln(A, lda, W, bias, X, ldx, X, M, N, K, alpha, row_valid, g2 ? 2 : 1, g1, b1, g2, b2, Y, ldy,
                      y_row_valid, eps, CFM_EPI_RESIDUAL, nullptr, 0, st);
  // unfused: residual GEMM then the standalone LayerNorm kernel (which needs contiguous rows)
  CFM_CHECK_ARG(ldx == N && ldy == N, "cfm_gemm_ln: the unfused path needs contiguous X and Y rows");
  int rc = cfm_gemm(A, lda, W, bias, X, ldx, M, N, K, dtype, CFM_EPI_RESIDUAL, X, alpha, row_valid,
                    engine == CFM_ENGINE_SIMT ? CFM_ENGINE_SIMT : CFM_ENGINE_AUTO, stream);
  if (rc != 0) return rc;
  return cfm_layernorm(X, M, N, g1, b1, g2 ? X : nullptr, g2, b2, Y, dtype, y_row_valid, eps, stream);
}

extern "C" int cfm_ffn(const void* y, int ld_in, const void* W1, const float* b1, const void* W2, const float* b2, float* X,
                       int ldx, int M, int d, int F, int dtype, float alpha, const float* g1, const float* be1,
                       const float* g2, const float* be2, void* Y, int ld_out, const uint8_t* y_row_valid, float eps,
                       void* hidden_ws, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(y && W1 && W2 && b1 && b2 && X, "cfm_ffn: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_ffn: bad dtype %d", dtype);
  CFM_CHECK_ARG((g1 == nullptr) == (be1 == nullptr) && (g2 == nullptr) == (be2 == nullptr) && (g1 || !g2),
                "cfm_ffn: inconsistent LayerNorm parameters");
  CFM_CHECK_ARG(g1 == nullptr || Y != nullptr, "cfm_ffn: LayerNorm requested but Y is null");
  CFM_CHECK_ARG(M >= 0 && d > 0 && F > 0 && ld_in >= d && ldx >= d, "cfm_ffn: bad shape");
  if (M == 0) return 0;
  const int ln_mode = g1 ? (g2 ? 2 : 1) : 0;
  const bool fused_ok = ffn_fused_supported(ld_in, ldx, ld_out, M, d, F, dtype, ln_mode);
  if (engine == CFM_ENGINE_TC)
    CFM_CHECK_ARG(fused_ok, "cfm_ffn: fused tcgen05 path does not support M=%d d=%d F=%d dtype=%d", M, d, F, dtype);
  if (fused_ok && engine != CFM_ENGINE_SIMT) {
    return ffn_fused(y, ld_in, W1, b1, W2, b2, X, ldx, M, F, alpha, ln_mode, g1, be1, g2, be2, Y, ld_out, y_row_valid,
                     eps, (cudaStream_t)stream);
  }
  CFM_CHECK_ARG(hidden_ws != nullptr, "cfm_ffn: the unfused path needs hidden_ws");
  int rc = cfm_gemm(y, ld_in, W1, b1, hidden_ws, F, M, F, d, dtype, CFM_EPI_BIAS_SILU, nullptr, 1.f, nullptr, engine, stream);
  if (rc != 0) return rc;
  if (ln_mode == 0)
    return cfm_gemm(hidden_ws, F, W2, b2, X, ldx, M, d, F, dtype, CFM_EPI_RESIDUAL, X, alpha, nullptr, engine, stream);
  return cfm_gemm_ln(hidden_ws, F, W2, b2, X, ldx, M, d, F, dtype, alpha, nullptr, g1, be1, g2, be2, Y, ld_out, y_row_valid,
                     eps, engine == CFM_ENGINE_TC ? CFM_ENGINE_AUTO : engine, stream);
}

extern "C" int cfm_mhsa_out(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts,
                            const void* v, int64_t v_bs, int64_t v_ts, int B, int H, int Tq, int Tk, const uint8_t* mask,
                            int64_t mask_bs, int64_t mask_rs, const float* key_bias, float scale, const void* Wo,
                            const float* bo, float* X, int dtype, const float* g1, const float* be1, void* Y,
                            const uint8_t* y_row_valid, float eps, void* ctx_ws, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(q && k && v && Wo && bo && X, "cfm_mhsa_out: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_mhsa_out: bad dtype %d", dtype);
  CFM_CHECK_ARG((g1 == nullptr) == (be1 == nullptr), "cfm_mhsa_out: inconsistent LayerNorm parameters");
  CFM_CHECK_ARG(g1 == nullptr || Y != nullptr, "cfm_mhsa_out: LayerNorm requested but Y is null");
  CFM_CHECK_ARG(B >= 0 && H > 0 && Tq >= 0 && Tk >= 0, "cfm_mhsa_out: bad shape");
  if (B == 0 || Tq == 0) return 0;
  const int d = H * 64, M = B * Tq;
  // CFM_B200_MHSA_MODE: "unfused" = always attention kernel + residual GEMM
  static const bool unfused_mode = env_is("CFM_B200_MHSA_MODE", "unfused");
  const bool fused_ok = scale > 0.f && mhsa_fused_supported(q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, B, H, Tq, Tk, d, dtype,
                                                            key_bias != nullptr);
  if (engine == CFM_ENGINE_TC)
    CFM_CHECK_ARG(fused_ok, "cfm_mhsa_out: fused tcgen05 path does not support H=%d Tq=%d Tk=%d dtype=%d", H, Tq, Tk, dtype);
  if (fused_ok && (engine == CFM_ENGINE_TC || (engine == CFM_ENGINE_AUTO && !unfused_mode)))
    return mhsa_fused(q, q_bs, q_ts, k, k_bs, k_ts, v, v_bs, v_ts, B, Tq, mask, mask_bs, mask_rs, scale, Wo, bo, X, g1, be1, Y,
                      y_row_valid, eps, (cudaStream_t)stream);
  CFM_CHECK_ARG(ctx_ws != nullptr, "cfm_mhsa_out: the unfused path needs ctx_ws");
  int rc = cfm_attention(q, q_bs, q_ts, k, k_bs, k_ts, v, v_bs, v_ts, ctx_ws, B, H, Tq, Tk, mask, mask_bs, mask_rs, key_bias,
                         scale, dtype, engine, stream);
  if (rc != 0) return rc;
  if (g1 == nullptr)
    return cfm_gemm(ctx_ws, d, Wo, bo, X, d, M, d, d, dtype, CFM_EPI_RESIDUAL, X, 1.f, nullptr, engine, stream);
  return cfm_gemm_ln(ctx_ws, d, Wo, bo, X, d, M, d, d, dtype, 1.f, nullptr, g1, be1, nullptr, nullptr, Y, d, y_row_valid, eps,
                     engine, stream);
}

extern "C" int cfm_conv_module(const void* y, const void* W1, const float* b1, const float* dw_w, const float* dw_b,
                               const void* W2, const float* b2, float* X, int B, int T, int d, int k, int dtype,
                               const uint8_t* row_valid, const float* g1, const float* be1, void* Y, float eps,
                               void* glu_ws, void* dw_ws, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(y && W1 && b1 && dw_w && dw_b && W2 && b2 && X, "cfm_conv_module: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_conv_module: bad dtype %d", dtype);
  CFM_CHECK_ARG((g1 == nullptr) == (be1 == nullptr), "cfm_conv_module: inconsistent LayerNorm parameters");
  CFM_CHECK_ARG(g1 == nullptr || (Y != nullptr && Y != y), "cfm_conv_module: LayerNorm output Y must be a buffer other than y");
  CFM_CHECK_ARG(B >= 0 && T >= 0 && d > 0 && k > 0 && (k & 1), "cfm_conv_module: bad shape");
  const int M = B * T;
  if (M == 0) return 0;
  // CFM_B200_CONV_MODE: "unfused" = always the three-kernel chain
  static const bool unfused_mode = env_is("CFM_B200_CONV_MODE", "unfused");
  const bool fused_ok = conv_fused_supported(M, T, d, k, dtype);
  if (engine == CFM_ENGINE_TC)
    CFM_CHECK_ARG(fused_ok, "cfm_conv_module: fused tcgen05 path does not support T=%d d=%d k=%d dtype=%d", T, d, k, dtype);
  if (fused_ok && (engine == CFM_ENGINE_TC || (engine == CFM_ENGINE_AUTO && !unfused_mode)))
    return conv_fused(y, W1, b1, dw_w, dw_b, W2, b2, X, M, T, row_valid, g1, be1, Y, eps, (cudaStream_t)stream);
  CFM_CHECK_ARG(glu_ws != nullptr && dw_ws != nullptr, "cfm_conv_module: the unfused path needs glu_ws and dw_ws");
  int rc = cfm_gemm(y, d, W1, b1, glu_ws, d, M, d, d, dtype, CFM_EPI_BIAS_GLU, nullptr, 1.f, nullptr, engine, stream);
  if (rc != 0) return rc;
  rc = cfm_dwconv(glu_ws, dw_w, dw_b, dw_ws, B, T, d, k, dtype, 1, stream);
  if (rc != 0) return rc;
  if (g1 == nullptr)
    return cfm_gemm(dw_ws, d, W2, b2, X, d, M, d, d, dtype, CFM_EPI_RESIDUAL, X, 1.f, row_valid, engine, stream);
  return cfm_gemm_ln(dw_ws, d, W2, b2, X, d, M, d, d, dtype, 1.f, row_valid, g1, be1, nullptr, nullptr, Y, d, nullptr, eps,
                     engine, stream);
}

extern "C" int cfm_ffn_chain(const void* y, int M, int d, int F, int dtype, const void* W1a, const float* b1a, const void* W2a,
                             const float* b2a, float alpha_a, const float* g1a, const float* be1a, const float* g2a,
                             const float* be2a, const void* W1b, const float* b1b, const void* W2b, const float* b2b,
                             float alpha_b, const float* g1b, const float* be1b, const float* g2b, const float* be2b, float* X,
                             void* Y, const uint8_t* y_row_valid, const void* Wp, const float* bp, void* P, int Np, float eps,
                             void* hidden_ws, int engine, void* stream) {
  using namespace cfm;
  const bool has_a = W1a != nullptr;
  CFM_CHECK_ARG(y && W1b && b1b && W2b && b2b && X && Y, "cfm_ffn_chain: null pointer");
  CFM_CHECK_ARG(!has_a || (b1a && W2a && b2a && g1a && be1a),
                "cfm_ffn_chain: the first module needs all its tensors and a LayerNorm (its output feeds the second)");
  CFM_CHECK_ARG((Wp == nullptr) || (bp && P && Np > 0 && g1b), "cfm_ffn_chain: projection needs bp, P, Np and a final LayerNorm");
  CFM_CHECK_ARG(M >= 0 && d > 0 && F > 0, "cfm_ffn_chain: bad shape");
  if (M == 0) return 0;
  // CFM_B200_FFN_CHAIN=0: always separate cfm_ffn / cfm_gemm calls
  static const bool chain_off = env_is("CFM_B200_FFN_CHAIN", "0");
  const FfnModule a{W1a, b1a, W2a, b2a, alpha_a, g1a, be1a, g2a, be2a};
  const FfnModule b{W1b, b1b, W2b, b2b, alpha_b, g1b, be1b, g2b, be2b};
  const bool ok = (dtype == CFM_BF16) && ffn_chain_supported(M, d, F, dtype, has_a ? &a : nullptr, b, Wp ? Np : 0);
  if (engine == CFM_ENGINE_TC) CFM_CHECK_ARG(ok, "cfm_ffn_chain: chained tcgen05 path does not support M=%d d=%d F=%d Np=%d", M, d, F, Np);
  if (ok && (engine == CFM_ENGINE_TC || (engine == CFM_ENGINE_AUTO && !chain_off)))
    return ffn_chain(y, has_a ? &a : nullptr, b, X, M, F, Y, y_row_valid, eps, Wp, bp, P, Np, (cudaStream_t)stream);
  int rc = 0;
  const void* yb = y;
  if (has_a) {
    rc = cfm_ffn(y, d, W1a, b1a, W2a, b2a, X, d, M, d, F, dtype, alpha_a, g1a, be1a, g2a, be2a, Y, d, nullptr, eps, hidden_ws,
                 engine, stream);
    if (rc != 0) return rc;
    yb = Y;
  }
  rc = cfm_ffn(yb, d, W1b, b1b, W2b, b2b, X, d, M, d, F, dtype, alpha_b, g1b, be1b, g2b, be2b, Y, d, y_row_valid, eps, hidden_ws,
               engine, stream);
  if (rc != 0 || Wp == nullptr) return rc;
  return cfm_gemm(Y, d, Wp, bp, P, Np, M, Np, d, dtype, CFM_EPI_BIAS, nullptr, 1.f, nullptr, engine, stream);
}

extern "C" int cfm_attention(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts,
                             const void* v, int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk,
                             const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, const float* key_bias,
                             float scale, int dtype, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(q && k && v && out, "cfm_attention: null q/k/v/out");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_attention: bad dtype %d", dtype);
  CFM_CHECK_ARG(B >= 0 && H > 0 && Tq >= 0 && Tk >= 0, "cfm_attention: bad shape");
  if (B == 0 || Tq == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = key_bias == nullptr && scale > 0.f &&
                     attention_tc_supported(q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, B, H, Tq, Tk, dtype);
  if (engine == CFM_ENGINE_TC) CFM_CHECK_ARG(tc_ok, "cfm_attention: tcgen05 engine does not support this shape/dtype");
  if (tc_ok && engine != CFM_ENGINE_SIMT) {
    // more than one 128-row query tile per (batch, head): the two-tile ping-pong kernel (CFM_B200_ATTN_PP=0 disables)
    static const bool pp_off = env_is("CFM_B200_ATTN_PP", "0");
    if (Tq > 128 && !pp_off)
      return attention_pp(q, q_bs, q_ts, k, k_bs, k_ts, v, v_bs, v_ts, out, B, H, Tq, Tk, mask, mask_bs, mask_rs, scale, st);
    return attention_tc(q, q_bs, q_ts, k, k_bs, k_ts, v, v_bs, v_ts, out, B, H, Tq, Tk, mask, mask_bs, mask_rs,
                        key_bias, scale, dtype, st);
  }
  return attention_simt(q, q_bs, q_ts, k, k_bs, k_ts, v, v_bs, v_ts, out, B, H, Tq, Tk, mask, mask_bs, mask_rs,
                        key_bias, scale, dtype, st);
}

extern "C" int cfm_gemm_ex(const void* A, int a_mn_major, int64_t lda, int64_t a_hs, int64_t a_bs, const void* B,
                           int b_mn_major, int64_t ldb, int64_t b_hs, int64_t b_bs, void* C, int c_dtype, int64_t ldc,
                           int64_t c_hs, int64_t c_bs, int accumulate, int M, int N, int K, int nH, int nB, int in_dtype,
                           float alpha, int splits, int engine, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(A && B && C, "cfm_gemm_ex: null A/B/C");
  CFM_CHECK_ARG(in_dtype == CFM_F32 || in_dtype == CFM_BF16, "cfm_gemm_ex: bad input dtype %d", in_dtype);
  CFM_CHECK_ARG(c_dtype == CFM_F32 || c_dtype == CFM_BF16, "cfm_gemm_ex: bad output dtype %d", c_dtype);
  CFM_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && nH >= 1 && nB >= 1, "cfm_gemm_ex: bad shape");
  CFM_CHECK_ARG(lda > 0 && ldb > 0 && ldc >= N, "cfm_gemm_ex: bad leading dimensions");
  if (M == 0 || N == 0) return 0;
  CFM_CHECK_ARG(K > 0, "cfm_gemm_ex: K must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = gemm_gen_tc_supported(A, lda, a_hs, a_bs, B, ldb, b_hs, b_bs, C, c_dtype, ldc, c_hs, c_bs, M, N, K, in_dtype);
  if (engine == CFM_ENGINE_TC)
    CFM_CHECK_ARG(tc_ok, "cfm_gemm_ex: tcgen05 engine does not support M=%d N=%d K=%d (dtype %d, strides/alignment)", M, N, K,
                  in_dtype);
  if (tc_ok && engine != CFM_ENGINE_SIMT)
    return gemm_gen(A, a_mn_major, lda, a_hs, a_bs, B, b_mn_major, ldb, b_hs, b_bs, C, c_dtype, ldc, c_hs, c_bs, accumulate, M, N,
                    K, nH, nB, alpha, splits, st);
  return gemm_gen_simt(A, a_mn_major, lda, a_hs, a_bs, B, b_mn_major, ldb, b_hs, b_bs, C, c_dtype, ldc, c_hs, c_bs, accumulate,
                       M, N, K, nH, nB, in_dtype, alpha, st);
}
