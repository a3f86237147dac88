// CUDA-core GEMM engine: C = epilogue(A W^T + bias) with fp32 accumulation.
// It is the engine of the fp32 path (bit-exact CTC-id gate) and of shapes the tcgen05
// engine does not take (tiny streaming chunks).  Classic 64x64x16 shared-memory tiling, 4x4
// register blocking; the epilogues are the ones listed in include/cfm_b200.h.
#include "cfm_common.cuh"

namespace cfm {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
}

template <typename T, int EPI>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, const float* __restrict__ bias,
                 void* Cv, int ldc, int M, int N, int K, const float* residual, float alpha,
                 const uint8_t* __restrict__ row_valid) {
  constexpr bool GLU = (EPI == CFM_EPI_BIAS_GLU);
  __shared__ __align__(16) float As[BK][BM + PADM];
  __shared__ __align__(16) float Ws[GLU ? 2 : 1][BK][BN + PADM];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;   // loader: row within tile, k offset

  float acc[4][4] = {}, acc2[GLU ? 4 : 1][4] = {};

  for (int k0 = 0; k0 < K; k0 += BK) {
    float va[4] = {0.f, 0.f, 0.f, 0.f}, vw[4] = {0.f, 0.f, 0.f, 0.f}, vw2[4] = {0.f, 0.f, 0.f, 0.f};
    const bool kin = (k0 + lk) < K;
    if (kin && (m0 + lr) < M) load4<T>(A + (size_t)(m0 + lr) * lda + k0 + lk, va);
    if (kin && (n0 + lr) < N) {
      load4<T>(W + (size_t)(n0 + lr) * K + k0 + lk, vw);
      if constexpr (GLU) load4<T>(W + (size_t)(N + n0 + lr) * K + k0 + lk, vw2);
    }
    __syncthreads();   // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = va[i];
      Ws[0][lk + i][lr] = vw[i];
      if constexpr (GLU) Ws[1][lk + i][lr] = vw2[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[0][kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, wr[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
      if constexpr (GLU) {
        const float4 w2 = *reinterpret_cast<const float4*>(&Ws[1][kk][tx * 4]);
        const float w2r[4] = {w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = fmaf(ar[i], w2r[j], acc2[i][j]);
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const bool valid = (row_valid == nullptr) || (row_valid[m] != 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      const size_t off = (size_t)m * ldc + n;
      if constexpr (EPI == CFM_EPI_BIAS) {
        ((T*)Cv)[off] = from_f32<T>(v);
      } else if constexpr (EPI == CFM_EPI_BIAS_SILU) {
        ((T*)Cv)[off] = from_f32<T>(act_silu<T>(v));
      } else if constexpr (EPI == CFM_EPI_BIAS_RELU) {
        ((T*)Cv)[off] = from_f32<T>(fmaxf(v, 0.f));
      } else if constexpr (EPI == CFM_EPI_BIAS_GLU) {
        const float g = acc2[i][j] + (bias ? bias[N + n] : 0.f);
        ((T*)Cv)[off] = from_f32<T>(v * act_sigmoid<T>(g));
      } else {
        if (!valid) v = 0.f;
        ((float*)Cv)[off] = (residual ? residual[off] : 0.f) + alpha * v;
      }
    }
  }
}

// ------------------------------------------------------------------ skinny GEMM: M <= 16 rows (streaming chunks)
// A streaming step (B = 1, 16 frames) is a chain of GEMMs with 16 rows: the tiled kernel above launches N/64 CTAs that
// walk K in 16-wide steps (4 CTAs x 128 synchronised steps for w_2) and takes 20-40 us per GEMM, although all there is
// to do is stream <= 1 MB of weights once.  Here a CTA owns two output columns and its eight warps split K four ways:
// every lane streams 8 consecutive k of its column's weight row per step (16 / 32 byte loads, fully coalesced across
// the warp) and of the 16 activation rows (L1 hits: all CTAs read the same <= 64 KB), accumulates 16 partial dot
// products in registers, the warp reduces them with shuffles and the four K-slices are summed through shared memory in
// a fixed order (deterministic).  N/2 CTAs (128 for the d = 256 outputs) keep the whole chip streaming.  fp32
// accumulation, same epilogues as above.
constexpr int SK_ROWS = 16, SK_WARPS = 8;

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&o)[8]) {
  if constexpr (sizeof(T) == 4) {
    const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
    o[0] = v0.x; o[1] = v0.y; o[2] = v0.z; o[3] = v0.w; o[4] = v1.x; o[5] = v1.y; o[6] = v1.z; o[7] = v1.w;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
  }
}

// SK_KS = warps that split K (4 for long K: 2 columns per CTA; 1 for short K: 8 columns per CTA)
template <typename T, int EPI, int SK_KS>
__global__ void __launch_bounds__(SK_WARPS * 32)
gemm_skinny_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, const float* __restrict__ bias, void* Cv,
                   int ldc, int M, int N, int K, const float* residual, float alpha,
                   const uint8_t* __restrict__ row_valid) {
  constexpr bool GLU = (EPI == CFM_EPI_BIAS_GLU);
  constexpr int kW = GLU ? 2 : 1;                       // weight rows per column (GLU: value + gate)
  constexpr int kColsCta = SK_WARPS / SK_KS;
  __shared__ float red[kColsCta][SK_KS][kW][SK_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = warp / SK_KS, ks = warp % SK_KS;       // column of the CTA, K slice
  const int n = min((int)blockIdx.x * kColsCta + cg, N - 1);
  float acc[kW][SK_ROWS];
#pragma unroll
  for (int c = 0; c < kW; ++c)
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) acc[c][r] = 0.f;
  for (int kk = (ks * 32 + lane) * 8; kk < K; kk += SK_KS * 256) {
    float w[kW][8];
#pragma unroll
    for (int c = 0; c < kW; ++c) load8<T>(W + (size_t)(c == 0 ? n : N + n) * K + kk, w[c]);
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) {
      float a[8];
      load8<T>(A + (size_t)min(r, M - 1) * lda + kk, a);
#pragma unroll
      for (int c = 0; c < kW; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[c][r] = fmaf(a[e], w[c][e], acc[c][r]);
    }
  }
#pragma unroll
  for (int c = 0; c < kW; ++c)
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r) {
      float v = acc[c][r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == r) red[cg][ks][c][r] = v;
    }
  __syncthreads();
  // warp (cg, 0): lane r < M finishes row r of column n
  if (ks != 0 || lane >= M || (int)blockIdx.x * kColsCta + cg >= N) return;
  const int m = lane;
  float v = 0.f, g = 0.f;
#pragma unroll
  for (int q = 0; q < SK_KS; ++q) {
    v += red[cg][q][0][m];
    if constexpr (GLU) g += red[cg][q][1][m];
  }
  v += bias ? bias[n] : 0.f;
  const size_t off = (size_t)m * ldc + n;
  if constexpr (EPI == CFM_EPI_BIAS) {
    ((T*)Cv)[off] = from_f32<T>(v);
  } else if constexpr (EPI == CFM_EPI_BIAS_SILU) {
    ((T*)Cv)[off] = from_f32<T>(act_silu<T>(v));
  } else if constexpr (EPI == CFM_EPI_BIAS_RELU) {
    ((T*)Cv)[off] = from_f32<T>(fmaxf(v, 0.f));
  } else if constexpr (EPI == CFM_EPI_BIAS_GLU) {
    g += bias ? bias[N + n] : 0.f;
    ((T*)Cv)[off] = from_f32<T>(v * act_sigmoid<T>(g));
  } else {
    const bool valid = (row_valid == nullptr) || (row_valid[m] != 0);
    if (!valid) v = 0.f;
    ((float*)Cv)[off] = (residual ? residual[off] : 0.f) + alpha * v;
  }
}

template <typename T, int EPI>
void launch_skinny(const T* a, int lda, const T* w, const float* bias, void* C, int ldc, int M, int N, int K,
                   const float* residual, float alpha, const uint8_t* rv, cudaStream_t st) {
  if (K >= 1024)
    gemm_skinny_kernel<T, EPI, 4><<<(N + 1) / 2, SK_WARPS * 32, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
  else
    gemm_skinny_kernel<T, EPI, 1><<<(N + 7) / 8, SK_WARPS * 32, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
}

template <typename T>
int launch(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N, int K,
           int epi, const float* residual, float alpha, const uint8_t* rv, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  const T* a = (const T*)A;
  const T* w = (const T*)W;
  if (M <= SK_ROWS && K % 8 == 0 && lda % 8 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(W) & 15) == 0) {     // streaming chunks: weight-streaming kernel
    switch (epi) {
      case CFM_EPI_BIAS: launch_skinny<T, CFM_EPI_BIAS>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv, st); break;
      case CFM_EPI_BIAS_SILU: launch_skinny<T, CFM_EPI_BIAS_SILU>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv, st); break;
      case CFM_EPI_BIAS_RELU: launch_skinny<T, CFM_EPI_BIAS_RELU>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv, st); break;
      case CFM_EPI_BIAS_GLU: launch_skinny<T, CFM_EPI_BIAS_GLU>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv, st); break;
      default: launch_skinny<T, CFM_EPI_RESIDUAL>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv, st); break;
    }
    CFM_LAUNCHED_K("gemm_skinny");
    return 0;
  }
  switch (epi) {
    case CFM_EPI_BIAS:
      gemm_simt_kernel<T, CFM_EPI_BIAS><<<grid, 256, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
      break;
    case CFM_EPI_BIAS_SILU:
      gemm_simt_kernel<T, CFM_EPI_BIAS_SILU><<<grid, 256, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
      break;
    case CFM_EPI_BIAS_RELU:
      gemm_simt_kernel<T, CFM_EPI_BIAS_RELU><<<grid, 256, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
      break;
    case CFM_EPI_BIAS_GLU:
      gemm_simt_kernel<T, CFM_EPI_BIAS_GLU><<<grid, 256, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
      break;
    default:
      gemm_simt_kernel<T, CFM_EPI_RESIDUAL><<<grid, 256, 0, st>>>(a, lda, w, bias, C, ldc, M, N, K, residual, alpha, rv);
      break;
  }
  CFM_LAUNCHED_K("gemm_simt");
  return 0;
}

}  // namespace

int gemm_simt(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N, int K,
              int dtype, int epilogue, const float* residual, float alpha, const uint8_t* row_valid,
              cudaStream_t st) {
  CFM_CHECK_ARG(K % 4 == 0 && lda % 4 == 0, "cfm_gemm(simt): K=%d and lda=%d must be multiples of 4", K, lda);
  CFM_CHECK_ARG((M + BM - 1) / BM <= 65535, "cfm_gemm(simt): M=%d too large", M);
  if (dtype == CFM_F32) return launch<float>(A, lda, W, bias, C, ldc, M, N, K, epilogue, residual, alpha, row_valid, st);
  return launch<__nv_bfloat16>(A, lda, W, bias, C, ldc, M, N, K, epilogue, residual, alpha, row_valid, st);
}

}  // namespace cfm
