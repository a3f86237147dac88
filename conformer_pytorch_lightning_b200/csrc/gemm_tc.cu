// tcgen05 GEMM engine for sm_100a:  C = epilogue(A W^T + bias), A (M,K) bf16, W (N,K) bf16 (both K-major),
// fp32 accumulation in TMEM.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0 (1 thread)  TMA producer: A tile 128x64 and W tile BNx64 (128-byte swizzle) into a
//                      STAGES-deep shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1 (1 thread)  MMA issuer: 4 x tcgen05.mma (M=128, N=BN, K=16) per stage into one of two TMEM
//                      accumulator buffers; tcgen05.commit releases the smem slot / publishes the tile
//   warp 2             TMEM allocator (2 x BN fp32 columns)
//   warps 4-7          epilogue: tcgen05.ld 32x32b (thread = output row), bias + SiLU / GLU / residual(+row
//                      mask) in registers, 16-byte global stores.  Runs concurrently with the main loop of
//                      the next tile thanks to the double-buffered accumulator.
// Replaces the nn.Linear / 1x1 Conv1d calls of feedforward.py:17-20, attention.py:62-64,99 and
// convolution.py:41-42,46 of the reference together with the elementwise ops that follow them.
#include "cfm_common.cuh"
#include "tc_common.cuh"

#include <mutex>

namespace cfm {
namespace tc {

// ------------------------------------------------------------------ host: driver entry point + tensor maps
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  CFM_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[3];
  cuuint64_t gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFM_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int BM = 128;       // rows per tile  = UMMA M
constexpr int BK = 64;        // K per stage    = one 128-byte swizzle atom of bf16
constexpr int UK = 16;        // K per tcgen05.mma (bf16)
constexpr int kThreads = 256;

template <int BN> struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  const float* bias;
  void* C;
  const float* residual;
  const uint8_t* row_valid;
  float alpha;
  int ldc, M, N, K;      // N = number of OUTPUT columns (GLU: W has 2N rows)
};

template <int EPI> __device__ __forceinline__ float epi_act(float v) {
  if constexpr (EPI == CFM_EPI_BIAS_SILU) return silu_fast(v);
  return v;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using C = Cfg<BN>;
  constexpr bool GLU = (EPI == CFM_EPI_BIAS_GLU);
  constexpr int OUT_BN = GLU ? BN / 2 : BN;      // output columns per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = p.N / OUT_BN;
  const int total = m_tiles * n_tiles;
  const int kb_count = p.K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0, phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int m0 = (t / n_tiles) * BM, nb = t % n_tiles;
      for (int kb = 0; kb < kb_count; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        uint8_t* sb = sa + C::kABytes;
        mbar_expect_tx(full_bar + stage, C::kStageBytes);
        tma_load_2d(sa, &tmA, full_bar + stage, kb * BK, m0);
        if constexpr (GLU) {   // value half and gate half of [Wa;Wb] side by side in one B tile
          tma_load_2d(sb, &tmW, full_bar + stage, kb * BK, nb * OUT_BN);
          tma_load_2d(sb + C::kBBytes / 2, &tmW, full_bar + stage, kb * BK, p.N + nb * OUT_BN);
        } else {
          tma_load_2d(sb, &tmW, full_bar + stage, kb * BK, nb * BN);
        }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar + acc, acc_phase ^ 1);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < kb_count; ++kb) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
        const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + C::kABytes);
#pragma unroll
        for (int k = 0; k < BK / UK; ++k) {
          // advance 32 bytes (16 bf16) inside the swizzle atom: +2 in the 16-byte-unit start address
          umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        umma_commit(empty_bar + stage);               // frees the smem slot when these MMAs retire
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar + acc);                   // accumulator complete -> epilogue
    }
  } else if (warp >= 4) {
    // ===================== epilogue (128 threads, thread = output row) =====================
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * OUT_BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      bool valid = true;
      if constexpr (EPI == CFM_EPI_RESIDUAL) valid = (p.row_valid == nullptr) || !row_ok || (p.row_valid[row] != 0);
#pragma unroll 1
      for (int c = 0; c < OUT_BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        [[maybe_unused]] uint32_t g[32];
        if constexpr (GLU) tmem_ld32(taddr + OUT_BN + c * 32, g);
        tmem_ld_wait();
        const int n = n0 + c * 32;
        if (row_ok) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
            f[j] = __uint_as_float(v[j]) + b4.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          if constexpr (GLU) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + p.N + n + j));
              f[j] *= sigmoid_fast(__uint_as_float(g[j]) + b4.x);
              f[j + 1] *= sigmoid_fast(__uint_as_float(g[j + 1]) + b4.y);
              f[j + 2] *= sigmoid_fast(__uint_as_float(g[j + 2]) + b4.z);
              f[j + 3] *= sigmoid_fast(__uint_as_float(g[j + 3]) + b4.w);
            }
          }
          if constexpr (EPI == CFM_EPI_RESIDUAL) {
            const size_t off = (size_t)row * p.ldc + n;
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + off);
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + off);
            const float a = valid ? p.alpha : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 r = r4[j];
              r.x = fmaf(a, f[4 * j], r.x);
              r.y = fmaf(a, f[4 * j + 1], r.y);
              r.z = fmaf(a, f[4 * j + 2], r.z);
              r.w = fmaf(a, f[4 * j + 3], r.w);
              o4[j] = r;
            }
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.ldc + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              o[j] = make_uint4(pack_bf16x2(epi_act<EPI>(f[8 * j]), epi_act<EPI>(f[8 * j + 1])),
                                pack_bf16x2(epi_act<EPI>(f[8 * j + 2]), epi_act<EPI>(f[8 * j + 3])),
                                pack_bf16x2(epi_act<EPI>(f[8 * j + 4]), epi_act<EPI>(f[8 * j + 5])),
                                pack_bf16x2(epi_act<EPI>(f[8 * j + 6]), epi_act<EPI>(f[8 * j + 7])));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::kTmemCols>(tmem_base);
}

template <int BN, int EPI>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CFM_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  constexpr int OUT_BN = (EPI == CFM_EPI_BIAS_GLU) ? BN / 2 : BN;
  const int total = ((p.M + BM - 1) / BM) * (p.N / OUT_BN);
  const int grid = total < num_sms() ? total : num_sms();
  gemm_tc_kernel<BN, EPI><<<grid, kThreads, C::kSmemBytes, st>>>(tmA, tmW, p);
  CFM_LAUNCHED();
  return 0;
}

// output-tile width: 256 when it divides N (fewer, fatter tiles), else 128
inline int pick_bn(int N, int epilogue) {
  if (epilogue == CFM_EPI_BIAS_GLU) return 256;          // 128 value + 128 gate columns
  return (N % 256 == 0) ? 256 : 128;
}

}  // namespace

bool gemm_tc_supported(int lda, int ldc, int M, int N, int K, int dtype, int epilogue) {
  if (dtype != CFM_BF16 || tc::encode_tiled_fn() == nullptr) return false;
  if (K % BK != 0 || lda % 8 != 0 || M < 1) return false;
  if (N % 128 != 0) return false;
  if (epilogue == CFM_EPI_RESIDUAL ? (ldc % 4 != 0) : (ldc % 8 != 0)) return false;
  if (M < 64) return false;     // tiny streaming chunks: a 128-row MMA tile is >50 % padding, SIMT engine is used
  return true;
}

int gemm_tc_init() {
  tc::encode_tiled_fn();
  return 0;
}

int gemm_tc(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N, int K, int dtype,
            int epilogue, const float* residual, float alpha, const uint8_t* row_valid, cudaStream_t st) {
  CFM_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                "cfm_gemm(tc): A/W/C must be 16-byte aligned");
  CFM_CHECK_ARG(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "cfm_gemm(tc): bias must be 16-byte aligned");
  const int bn = pick_bn(N, epilogue);
  const int w_rows = (epilogue == CFM_EPI_BIAS_GLU) ? 2 * N : N;
  CUtensorMap tmA, tmW;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t str[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {BK, BM};
    int rc = tc::make_tmap_bf16(&tmA, A, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)w_rows};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {BK, (uint32_t)((epilogue == CFM_EPI_BIAS_GLU) ? 128 : bn)};
    int rc = tc::make_tmap_bf16(&tmW, W, 2, dims, str, box);
    if (rc) return rc;
  }
  GemmParams p{bias, C, residual, row_valid, alpha, ldc, M, N, K};
  if (bn == 256) {
    switch (epilogue) {
      case CFM_EPI_BIAS: return launch_tc<256, CFM_EPI_BIAS>(tmA, tmW, p, st);
      case CFM_EPI_BIAS_SILU: return launch_tc<256, CFM_EPI_BIAS_SILU>(tmA, tmW, p, st);
      case CFM_EPI_BIAS_GLU: return launch_tc<256, CFM_EPI_BIAS_GLU>(tmA, tmW, p, st);
      default: return launch_tc<256, CFM_EPI_RESIDUAL>(tmA, tmW, p, st);
    }
  }
  switch (epilogue) {
    case CFM_EPI_BIAS: return launch_tc<128, CFM_EPI_BIAS>(tmA, tmW, p, st);
    case CFM_EPI_BIAS_SILU: return launch_tc<128, CFM_EPI_BIAS_SILU>(tmA, tmW, p, st);
    default: return launch_tc<128, CFM_EPI_RESIDUAL>(tmA, tmW, p, st);
  }
}

}  // namespace cfm
