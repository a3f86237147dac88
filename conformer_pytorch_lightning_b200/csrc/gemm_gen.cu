// General (transposed-operand, batched, split-K) tcgen05 GEMM for the TRAINING path:
//
//     C[b][h] (M x N)  (+)=  alpha * opA(A[b][h]) (M x K) * opB(B[b][h])^T (N x K)
//
// Each operand is either K-major (element (mn,k) at base + mn*ld + k: what nn.Linear's forward uses) or MN-major
// (element (mn,k) at base + k*ld + mn): with the two flags one kernel covers
//   dgrad   dA = dC W            A = dC  K-major,  B = W   MN-major      (autograd of F.linear, feedforward.py:17-21 ...)
//   wgrad   dW = dC^T A          A = dC  MN-major, B = Act MN-major, fp32 accumulate into the gradient, split over K = tokens
//   scores  S  = Q K^T           both K-major, batched over (batch, head)               (attention.py:84)
//   context O  = P V             A = P K-major, B = V MN-major                           (attention.py:96)
//   and the four batched products of the attention backward (dP = dO V^T, dV = P^T dO, dQ = dS K, dK = dS^T Q).
// MN-major operands are read by tcgen05.mma straight from 128-byte-swizzled [64 k-rows x 64 mn] TMA boxes (UMMA
// "MN-major" smem descriptors: leading byte offset = distance between 64-element atoms along MN, stride byte offset =
// 1024 B between 8-row k groups), so no transposed copy of any activation or weight is ever made.
//
// Kernel: persistent, warp-specialised like gemm_tc.cu -- warp 0 TMA producer (4-D tensor maps: inner, rows, head,
// batch), warp 1 MMA issuer (M=128, N=BN in {64,128,256}, K=16, two TMEM accumulators), warp 2 TMEM allocator, warps
// 4-7 epilogue (thread = output row; staging tiles leave through TMA stores, or TMA *reduce-add* stores for fp32
// accumulation, which is also what makes split-K race free).
// A CUDA-core kernel with the same semantics (any strides, fp32 or bf16) is the fp32 parity engine and the fallback.
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include "resid_epilogue.cuh"

namespace cfm {
namespace {

using namespace tc;

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int kAtomBytes = 64 * 128;      // one [64 rows x 128 B] swizzle-atom column of an MN-major operand

struct GenParams {
  int M, N, K, nH, nB;
  int m_tiles, n_tiles, splits, kb_total;
  float alpha;
  int out_mode;                            // 0: bf16 store, 1: fp32 store, 2: fp32 reduce-add
};

template <int BN> struct GCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 5 : 6);
  static constexpr int kBufs = 2;
  static constexpr int kThreads = 256;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBufs * kBufBytes + 256 + 1024;
  static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// UMMA shared-memory descriptor of an MN-major, 128-byte-swizzled operand made of [64 k x 64 mn] atoms kAtomBytes apart
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(kAtomBytes >> 4) << 16) |
         (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_gen(int m, int n, int a_mn, int b_mn) {
  return umma_idesc_bf16(m, n, b_mn) | (static_cast<uint32_t>(a_mn) << 15);
}

struct TileCoord { int b, h, m0, n0, kb0, kb1; };
__device__ __forceinline__ TileCoord decode_tile(int t, const GenParams& p, int BN) {
  TileCoord c;
  const int split = t % p.splits; t /= p.splits;
  c.n0 = (t % p.n_tiles) * BN; t /= p.n_tiles;
  c.m0 = (t % p.m_tiles) * BM; t /= p.m_tiles;
  c.h = t % p.nH;
  c.b = t / p.nH;
  c.kb0 = (int)((long long)split * p.kb_total / p.splits);
  c.kb1 = (int)((long long)(split + 1) * p.kb_total / p.splits);
  return c;
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_gen_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GenParams p) {
  using C = GCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + C::kStages * C::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + C::kBufs * kBufBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.nB * p.nH * p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<C::kTmemCols>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0, phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const TileCoord c = decode_tile(t, p, BN);
      for (int kb = c.kb0; kb < c.kb1; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          mbar_expect_tx(full_bar + stage, C::kStageBytes);
          if constexpr (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_4d(sa + j * kAtomBytes, &tmA, full_bar + stage, c.m0 + 64 * j, kb * BK, c.h, c.b);
          } else {
            tma_load_4d(sa, &tmA, full_bar + stage, kb * BK, c.m0, c.h, c.b);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * kAtomBytes, &tmB, full_bar + stage, c.n0 + 64 * j, kb * BK, c.h, c.b);
          } else {
            tma_load_4d(sb, &tmB, full_bar + stage, kb * BK, c.n0, c.h, c.b);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_gen(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const TileCoord c = decode_tile(t, p, BN);
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = c.kb0; kb < c.kb1; ++kb) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t da = A_MN ? umma_desc_mn_sw128(sa) : umma_desc_sw128(sa);
          const uint64_t db = B_MN ? umma_desc_mn_sw128(sa + C::kABytes) : umma_desc_sw128(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // K-major: 32 bytes per K step inside the swizzle atom; MN-major: 16 k-rows of 128 bytes = 2048 bytes
            umma_bf16(tmem_d, da + (A_MN ? 128 * k : 2 * k), db + (B_MN ? 128 * k : 2 * k), idesc, (kb != c.kb0 || k != 0));
          }
          umma_commit(empty_bar + stage);
          if (kb == c.kb1 - 1) umma_commit(tfull_bar + acc);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: thread = output row =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool elected = (threadIdx.x == 128);
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int it = 0, sub_cnt = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const TileCoord c = decode_tile(t, p, BN);
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + lane_base + acc * BN;
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
      if (p.out_mode == 0) {
        // bf16: 64 columns per staging tile
#pragma unroll 1
        for (int sub = 0; sub < BN / 64; ++sub, ++sub_cnt) {
          uint8_t* buf = ring + (sub_cnt & 1) * kBufBytes;
          if (elected) bulk_wait_read<1>();
          named_bar_sync(1, 128);
          uint32_t v[64];
          {
            uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
            uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
            tmem_ld32(taddr + sub * 64, v0);
            tmem_ld32(taddr + sub * 64 + 32, v1);
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = p.alpha * __uint_as_float(v[8 * j + e]);
            *reinterpret_cast<uint4*>(buf + sw_off(r, j)) =
                make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (elected) {
            if (c.n0 + sub * 64 < p.N) tma_store_4d(&tmC, buf, c.n0 + sub * 64, c.m0, c.h, c.b);
            bulk_commit();
          }
        }
      } else {
        // fp32: 32 columns per staging tile; plain store or reduce-add
#pragma unroll 1
        for (int sub = 0; sub < BN / 32; ++sub, ++sub_cnt) {
          uint8_t* buf = ring + (sub_cnt & 1) * kBufBytes;
          if (elected) bulk_wait_read<1>();
          named_bar_sync(1, 128);
          uint32_t v[32];
          tmem_ld32(taddr + sub * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<float4*>(buf + sw_off(r, j)) =
                make_float4(p.alpha * __uint_as_float(v[4 * j]), p.alpha * __uint_as_float(v[4 * j + 1]),
                            p.alpha * __uint_as_float(v[4 * j + 2]), p.alpha * __uint_as_float(v[4 * j + 3]));
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (elected) {
            if (c.n0 + sub * 32 < p.N) {
              if (p.out_mode == 2) tma_reduce_add_4d(&tmC, buf, c.n0 + sub * 32, c.m0, c.h, c.b);
              else tma_store_4d(&tmC, buf, c.n0 + sub * 32, c.m0, c.h, c.b);
            }
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + acc);
    }
    if (elected) bulk_wait_all<0>();      // reductions must have been performed before the grid is considered complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<C::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------ CUDA-core engine (parity / fallback)
template <typename TI, typename TO, bool ACC>
__global__ void __launch_bounds__(256)
gemm_gen_simt_kernel(const TI* __restrict__ A, long long a_ms, long long a_ks, long long a_hs, long long a_bs,
                     const TI* __restrict__ B, long long b_ns, long long b_ks, long long b_hs, long long b_bs, TO* C,
                     long long ldc, long long c_hs, long long c_bs, int M, int N, int K, int nH, float alpha) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int z = blockIdx.z, h = z % nH, b = z / nH;
  A += (long long)h * a_hs + (long long)b * a_bs;
  B += (long long)h * b_hs + (long long)b * b_bs;
  C += (long long)h * c_hs + (long long)b * c_bs;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      // choose the index split so that the unit-stride dimension varies fastest across threads
      int mm, kk;
      if (a_ks == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? to_f32(A[(long long)m * a_ms + (long long)k * a_ks]) : 0.f;
      int nn;
      if (b_ks == 1) { kk = i & 15; nn = i >> 4; } else { nn = i & 63; kk = i >> 6; }
      const int n = n0 + nn;
      const int k2 = k0 + kk;
      Bs[kk][nn] = (n < N && k2 < K) ? to_f32(B[(long long)n * b_ns + (long long)k2 * b_ks]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; bb[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      TO* dst = C + (long long)m * ldc + n;
      float v = alpha * acc[i][j];
      if constexpr (ACC) v += to_f32(*dst);
      *dst = from_f32<TO>(v);
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
int launch_gen(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GenParams& p, cudaStream_t st) {
  using C = GCfg<BN>;
  CFM_SMEM_OPT_IN((gemm_gen_kernel<BN, A_MN, B_MN>), C::kSmemBytes);
  const int total = p.nB * p.nH * p.m_tiles * p.n_tiles * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  CFM_CUDA_OK(launch_pdl(gemm_gen_kernel<BN, A_MN, B_MN>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st, 1, tmA, tmB, tmC, p));
  CFM_LAUNCHED_K("gemm_gen");
  return 0;
}

template <int BN>
int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                   const GenParams& p, cudaStream_t st) {
  if (a_mn) return b_mn ? launch_gen<BN, true, true>(tmA, tmB, tmC, p, st) : launch_gen<BN, true, false>(tmA, tmB, tmC, p, st);
  return b_mn ? launch_gen<BN, false, true>(tmA, tmB, tmC, p, st) : launch_gen<BN, false, false>(tmA, tmB, tmC, p, st);
}

// operand tensor map: dims innermost first.  K-major: (K, MN, H, B), box (64, rows);  MN-major: (MN, K, H, B), box (64, 64)
int operand_map(CUtensorMap* tm, const void* base, bool mn_major, int MN, int K, long long ld, long long hs, long long bs,
                int nH, int nB, int box_rows) {
  const uint64_t dims[4] = {(uint64_t)(mn_major ? MN : K), (uint64_t)(mn_major ? K : MN), (uint64_t)nH, (uint64_t)nB};
  // a stride of 0 is not encodable: single-entry dims get any legal (16-byte multiple) stride
  const uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)(nH > 1 ? hs : ld) * 2, (uint64_t)(nB > 1 ? bs : ld) * 2};
  const uint32_t box[4] = {64, (uint32_t)(mn_major ? 64 : box_rows), 1, 1};
  return tc::make_tmap_bf16(tm, base, 4, dims, str, box);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool gemm_gen_tc_supported(const void* A, long long lda, long long a_hs, long long a_bs, const void* B, long long ldb,
                           long long b_hs, long long b_bs, const void* C, int c_dtype, long long ldc, long long c_hs,
                           long long c_bs, int M, int N, int K, int in_dtype) {
  if (in_dtype != CFM_BF16 || tc::encode_tiled_fn() == nullptr) return false;
  if (!aligned16(A) || !aligned16(B) || !aligned16(C)) return false;
  if ((lda | a_hs | a_bs | ldb | b_hs | b_bs) % 8 != 0) return false;
  const int cq = (c_dtype == CFM_F32) ? 4 : 8;
  if ((ldc | c_hs | c_bs) % cq != 0) return false;
  if (M < 32 || N < 16 || K < 16) return false;
  return true;
}

int gemm_gen(const void* A, int a_mn, long long lda, long long a_hs, long long a_bs, const void* B, int b_mn, long long ldb,
             long long b_hs, long long b_bs, void* C, int c_dtype, long long ldc, long long c_hs, long long c_bs,
             int accumulate, int M, int N, int K, int nH, int nB, float alpha, int splits, cudaStream_t st) {
  int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  if (!(c_dtype == CFM_F32 && accumulate)) {
    // no split-K for a bf16 / overwritten output: narrower tiles instead when 256-wide ones leave most SMs idle (input
    // gradients at the C5 shard: 3968 x 256 with K = 2048 is 31 tiles of 128 x 256 -- 8.5 us of MMA time on 31 SMs -- or
    // 124 tiles of 128 x 64)
    const long long mt = (long long)nB * nH * ((M + BM - 1) / BM);
    while (bn > 64 && mt * ((N + bn - 1) / bn) * 2 <= num_sms()) bn /= 2;
  }
  GenParams p{};
  p.M = M; p.N = N; p.K = K; p.nH = nH; p.nB = nB;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + bn - 1) / bn;
  p.kb_total = (K + BK - 1) / BK;
  p.alpha = alpha;
  p.out_mode = (c_dtype == CFM_BF16) ? 0 : (accumulate ? 2 : 1);
  CFM_CHECK_ARG(!(accumulate && c_dtype != CFM_F32), "cfm_gemm_ex(tc): accumulation needs an fp32 output");
  const int tiles = p.nB * p.nH * p.m_tiles * p.n_tiles;
  if (p.out_mode != 2) splits = 1;
  else if (splits <= 0) {
    // ONE wave of (tile, K-slice) work items: every extra slice costs a full fp32 reduce-add of the output tile
    // (128 KB at ~31 B/clk/SM), so slices are only added until the chip is full, and keep >= 2 K-steps of 64 each
    splits = num_sms() / tiles;
    if (splits > p.kb_total / 2) splits = p.kb_total / 2;
  }
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.splits = splits;
  CUtensorMap tmA, tmB, tmC;
  int rc;
  if ((rc = operand_map(&tmA, A, a_mn != 0, M, K, lda, a_hs, a_bs, nH, nB, BM)) != 0) return rc;
  if ((rc = operand_map(&tmB, B, b_mn != 0, N, K, ldb, b_hs, b_bs, nH, nB, bn)) != 0) return rc;
  {
    const bool f32 = c_dtype == CFM_F32;
    const int es = f32 ? 4 : 2;
    const uint64_t dims[4] = {(uint64_t)N, (uint64_t)M, (uint64_t)nH, (uint64_t)nB};
    const uint64_t str[3] = {(uint64_t)ldc * es, (uint64_t)(nH > 1 ? c_hs : ldc) * es, (uint64_t)(nB > 1 ? c_bs : ldc) * es};
    const uint32_t box[4] = {(uint32_t)(f32 ? 32 : 64), (uint32_t)BM, 1, 1};
    rc = f32 ? tc::make_tmap_f32(&tmC, C, 4, dims, str, box) : tc::make_tmap_bf16(&tmC, C, 4, dims, str, box);
    if (rc != 0) return rc;
  }
  if (bn == 256) return dispatch_major<256>(a_mn != 0, b_mn != 0, tmA, tmB, tmC, p, st);
  if (bn == 128) return dispatch_major<128>(a_mn != 0, b_mn != 0, tmA, tmB, tmC, p, st);
  return dispatch_major<64>(a_mn != 0, b_mn != 0, tmA, tmB, tmC, p, st);
}

int gemm_gen_simt(const void* A, int a_mn, long long lda, long long a_hs, long long a_bs, const void* B, int b_mn,
                  long long ldb, long long b_hs, long long b_bs, void* C, int c_dtype, long long ldc, long long c_hs,
                  long long c_bs, int accumulate, int M, int N, int K, int nH, int nB, int in_dtype, float alpha,
                  cudaStream_t st) {
  const long long a_ms = a_mn ? 1 : lda, a_ks = a_mn ? lda : 1;
  const long long b_ns = b_mn ? 1 : ldb, b_ks = b_mn ? ldb : 1;
  dim3 grid((N + 63) / 64, (M + 63) / 64, nH * nB);
  CFM_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "cfm_gemm_ex(simt): shape too large");
#define CFM_GEN_SIMT(TI, TO, ACC)                                                                                       \
  gemm_gen_simt_kernel<TI, TO, ACC><<<grid, 256, 0, st>>>((const TI*)A, a_ms, a_ks, a_hs, a_bs, (const TI*)B, b_ns, b_ks, \
                                                          b_hs, b_bs, (TO*)C, ldc, c_hs, c_bs, M, N, K, nH, alpha)
  if (in_dtype == CFM_F32) {
    CFM_CHECK_ARG(c_dtype == CFM_F32, "cfm_gemm_ex(simt): fp32 inputs need an fp32 output");
    if (accumulate) CFM_GEN_SIMT(float, float, true); else CFM_GEN_SIMT(float, float, false);
  } else if (c_dtype == CFM_F32) {
    if (accumulate) CFM_GEN_SIMT(__nv_bfloat16, float, true); else CFM_GEN_SIMT(__nv_bfloat16, float, false);
  } else {
    CFM_CHECK_ARG(!accumulate, "cfm_gemm_ex(simt): accumulation needs an fp32 output");
    CFM_GEN_SIMT(__nv_bfloat16, __nv_bfloat16, false);
  }
#undef CFM_GEN_SIMT
  CFM_LAUNCHED_K("gemm_gen_simt");
  return 0;
}

}  // namespace cfm
