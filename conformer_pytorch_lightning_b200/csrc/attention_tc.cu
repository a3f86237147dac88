// tcgen05 flash attention for sm_100a (head dim 64, bf16 operands, fp32 softmax statistics / accumulation).
//
// One CTA = 128 query rows of one (batch, head); loop over 128-key tiles:
//   warp 4 (1 thread)  TMA producer + MMA issuer:  S = Q K^T (tcgen05.mma M=128 N=128 K=16 x4, both operands
//                      K-major, 128-byte swizzle) into TMEM cols [0,128);  O_tile = P V (M=128 N=64 K=16 x8, A = P
//                      from shared memory, B = V tile straight from TMA used as an MN-major operand) into TMEM
//                      cols [128,192).  K/V of the next tile are prefetched as soon as the MMA that reads the
//                      current one has retired (tcgen05.commit -> mbarrier).
//   warps 0-3          softmax: thread = query row.  Two passes of tcgen05.ld over the S row (max, then exp2),
//                      mask bytes -> -inf exactly like attention.py:89-92 of the reference (masked probabilities are
//                      0, a fully masked row outputs 0), P written as bf16 into the swizzled A-operand layout,
//                      running (max, sum) and the 64-wide output row kept in registers and rescaled online.
//   warp 5             TMEM allocator.
// Two CTAs fit per SM (80 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
// No T x T tensor is ever materialised (the reference creates ~8 of them, attention.py:84-96).
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace cfm {
namespace {

using namespace tc;

constexpr int QT = 128, KT = 128, DK = 64;
constexpr int kThreads = 192;
constexpr int kTileBytes = 128 * DK * 2;       // 16 KB: Q, K or V tile
constexpr int kPBytes = QT * KT * 2;           // 32 KB
constexpr int kSmemBytes = 3 * kTileBytes + kPBytes + 1024 + 128;
constexpr int kTmemCols = 256;

struct AttnParams {
  __nv_bfloat16* out;
  const uint8_t* mask;
  int64_t mask_bs, mask_rs;
  int H, Tq, Tk;
  float scale_log2;     // softmax scale * log2(e)
  int mask_aligned8;
  long long* trace;   // optional clock64 timeline of CTA (0,0,0) (tools/attn_trace.py); nullptr in production
};

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 32 mask bytes -> bitmask (bit c set = visible)
__device__ __forceinline__ uint32_t mask_bits32(const uint8_t* p, bool aligned8) {
  uint32_t bits = 0;
  if (aligned8) {
    const uint2* p2 = reinterpret_cast<const uint2*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint2 w = __ldg(p2 + i);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        bits |= (((w.x >> (8 * b)) & 0xffu) != 0u ? 1u : 0u) << (8 * i + b);
        bits |= (((w.y >> (8 * b)) & 0xffu) != 0u ? 1u : 0u) << (8 * i + 4 + b);
      }
    }
  } else {
#pragma unroll 8
    for (int c = 0; c < 32; ++c) bits |= (__ldg(p + c) != 0 ? 1u : 0u) << c;
  }
  return bits;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  // 1024-byte alignment is what SWIZZLE_128B tiles need; keeping `smem` a plain shared-space array (no integer
  // round-up) lets ptxas emit LDS/STS instead of generic LD.E/ST.E for every epilogue access
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 2 * kTileBytes;
  uint8_t* sP = smem + 3 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTileBytes + kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* s_full = bars + 3;    // S = QK^T landed in TMEM (also: K tile free)
  uint64_t* p_ready = bars + 4;   // P written to smem by the 128 softmax threads (also: S and O_tile drained)
  uint64_t* o_full = bars + 5;    // O_tile = PV landed in TMEM (also: P and V tile free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  uint32_t* svis = tmem_slot + 1;   // [2][4] visibility words of a broadcast (B,1,Tk) mask, double buffered by tile parity

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * QT;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Tk + KT - 1) / KT;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define ATR(slot) do { if (tr && lane == 0) p.trace[(slot)] = clock64(); } while (0)

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(v_full, 1);
    mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                   // nothing above touches global memory produced by the previous kernel
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;

  if (warp == 4) {
    // whole warp converged; one elected lane issues the async instructions (see elect_one in tc_common.cuh)
    constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KT, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(QT, DK, 1);   // B = V tile, MN-major (dk contiguous)
    if (elect_one()) {
      mbar_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &tmQ, q_full, h * DK, i0, b);
      mbar_expect_tx(k_full, kTileBytes);
      tma_load_3d(sK, &tmK, k_full, h * DK, 0, b);
      mbar_expect_tx(v_full, kTileBytes);
      tma_load_3d(sV, &tmV, v_full, h * DK, 0, b);
    }
    __syncwarp();
    ATR(0);
    mbar_wait(q_full, 0);
    ATR(1);
    const uint64_t dq = umma_desc_sw128(smem_u32(sQ)), dk = umma_desc_sw128(smem_u32(sK));
    const uint64_t dv = umma_desc_sw128(smem_u32(sV));
    const uint64_t dp0 = umma_desc_sw128(smem_u32(sP)), dp1 = umma_desc_sw128(smem_u32(sP) + kPBytes / 2);
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(k_full, ph);
      tc_fence_after();
      ATR(8 + j * 8 + 0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      __syncwarp();
      ATR(8 + j * 8 + 1);
      mbar_wait(s_full, ph);
      ATR(8 + j * 8 + 2);                       // S done -> K tile may be overwritten
      if (j + 1 < n_kv) {
        if (elect_one()) {
          mbar_expect_tx(k_full, kTileBytes);
          tma_load_3d(sK, &tmK, k_full, h * DK, (j + 1) * KT, b);
        }
        __syncwarp();
      }
      mbar_wait(p_ready, ph);                      // P in smem, S / O_tile drained by the softmax warps
      ATR(8 + j * 8 + 3);
      mbar_wait(v_full, ph);
      tc_fence_after();
      ATR(8 + j * 8 + 4);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          // A: P, K-major, two 64-key swizzle atoms (16 KB each), 32 bytes per K step inside an atom
          const uint64_t da = ((k < 4) ? dp0 : dp1) + 2 * (k & 3);
          // B: V tile rows = keys (128 bytes each): 16 keys per K step = 2048 bytes = +128 in 16-byte units
          umma_bf16(tmem_o, da, dv + 128 * k, idesc_o, k != 0);
        }
        umma_commit(o_full);
      }
      __syncwarp();
      ATR(8 + j * 8 + 5);
      mbar_wait(o_full, ph);                       // PV done -> V tile and P may be overwritten
      ATR(8 + j * 8 + 6);
      if (j + 1 < n_kv) {
        if (elect_one()) {
          mbar_expect_tx(v_full, kTileBytes);
          tma_load_3d(sV, &tmV, v_full, h * DK, (j + 1) * KT, b);
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ===================== softmax / accumulate / store: thread = query row =====================
    const int r = warp * 32 + lane;
    const int i = i0 + r;
    const bool row_ok = i < p.Tq;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint8_t* mrow = (p.mask != nullptr && row_ok) ? p.mask + b * p.mask_bs + i * p.mask_rs : nullptr;
    float m_run = -CUDART_INF_F, l_run = 0.f;
    float o[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) o[c] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int j0 = j * KT;
      // visibility bits of this row for the 4 x 32 keys of the tile
      uint32_t vis[4];
      if (p.mask != nullptr && p.mask_rs == 0) {
        // (B,1,Tk) key-padding mask: the 128 threads fetch one byte each and ballot -> 4 words shared by all rows
        const int jj = j0 + r;
        const bool on = (jj < p.Tk) && (__ldg(p.mask + b * p.mask_bs + jj) != 0);
        const uint32_t w = __ballot_sync(0xffffffffu, on);
        if (lane == 0) svis[(j & 1) * 4 + warp] = w;
        named_bar_sync(1, 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) vis[c] = row_ok ? svis[(j & 1) * 4 + c] : 0u;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jj = j0 + c * 32;
          const int nvalid = p.Tk - jj;
          uint32_t bits = !row_ok ? 0u : (nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u)));
          if (mrow != nullptr && bits != 0u) {
            if (nvalid >= 32) bits &= mask_bits32(mrow + jj, p.mask_aligned8 != 0);
            else {
              uint32_t mb = 0;
              for (int c2 = 0; c2 < nvalid; ++c2) mb |= (__ldg(mrow + jj + c2) != 0 ? 1u : 0u) << c2;
              bits &= mb;
            }
          }
          vis[c] = bits;
        }
      }
      if (warp == 0) ATR(64 + j * 8 + 0);
      mbar_wait(s_full, ph);
      tc_fence_after();
      if (warp == 0) ATR(64 + j * 8 + 1);
      // pass 1: row max over visible keys.  Fast path per 32-key chunk when every key is visible (the common case:
      // masking only bites on the tail tile / chunk boundaries) -- 1 FMNMX per element instead of shift+and+setp+sel.
      float m_tile = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_base + c * 32, v);
        tmem_ld_wait();
        if (vis[c] == 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) m_tile = fmaxf(m_tile, __uint_as_float(v[e]));
        } else if (vis[c] != 0u) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((vis[c] >> e) & 1u) m_tile = fmaxf(m_tile, __uint_as_float(v[e]));
        }
      }
      if (warp == 0) ATR(64 + j * 8 + 2);
      m_tile *= p.scale_log2;                               // scale > 0: max commutes with the scaling
      const float m_new = fmaxf(m_run, m_tile);
      const bool any = m_new != -CUDART_INF_F;
      const float alpha = any ? exp2f(m_run - m_new) : 1.f; // m_run = -inf, m_new finite -> 0
      const float neg_m = any ? -m_new : 0.f;
      // pass 2: p = 2^(s*c - m) (one FFMA + one MUFU.EX2 per element; masked keys are turned into -inf first so
      // they come out as exactly 0), P -> smem as bf16 in the swizzled K-major A-operand layout, row sum
      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_base + c * 32, v);
        tmem_ld_wait();
        if (vis[c] != 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((vis[c] >> e) & 1u)) v[e] = 0xff800000u;  // -inf
        }
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ex2_fast(fmaf(__uint_as_float(v[e]), p.scale_log2, neg_m));
          const float p1 = ex2_fast(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, neg_m));
          l_tile += p0 + p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        // 4 x 16-byte chunks of this row: global chunk index cc = 4c+q in [0,16); atom = cc/8, column cw = cc%8
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = 4 * c + q;
          uint8_t* dst = sP + (cc >> 3) * (kPBytes / 2) + r * 128 + (((cc & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      fence_proxy_async_smem();        // P stores -> visible to the tensor core (async proxy)
      tc_fence_before();               // our TMEM loads of S are complete (wait::ld above)
      mbar_arrive(p_ready);
      if (warp == 0) ATR(64 + j * 8 + 3);
      // O_tile = P V of this tile, accumulate into registers with the online rescale
      mbar_wait(o_full, ph);
      tc_fence_after();
      if (warp == 0) ATR(64 + j * 8 + 4);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_o + lane_base + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[c * 32 + e] = fmaf(o[c * 32 + e], alpha, __uint_as_float(v[e]));
      }
      tc_fence_before();
      if (warp == 0) ATR(64 + j * 8 + 5);
    }
    if (row_ok) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;     // fully masked row -> 0 (attention.py:92)
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.Tq + i) * p.H * DK + h * DK);
#pragma unroll
      for (int c = 0; c < DK / 8; ++c)
        dst[c] = make_uint4(pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                            pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<kTmemCols>(tmem_base);
}

int make_qkv_tmap(CUtensorMap* tm, const void* base, int B, int T, int H, int64_t bs, int64_t ts) {
  const uint64_t dims[3] = {(uint64_t)H * DK, (uint64_t)T, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)ts * 2, (uint64_t)bs * 2};
  const uint32_t box[3] = {DK, 128, 1};
  return tc::make_tmap_bf16(tm, base, 3, dims, str, box);
}

}  // namespace

bool attention_tc_supported(int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts, int B,
                            int H, int Tq, int Tk, int dtype) {
  if (dtype != CFM_BF16 || tc::encode_tiled_fn() == nullptr) return false;
  if (Tq < 32 || Tk < 1) return false;           // tiny streaming chunks stay on the CUDA-core kernel
  if ((q_bs | q_ts | k_bs | k_ts | v_bs | v_ts) % 8 != 0) return false;
  if (B > 65535 || H > 65535) return false;
  return true;
}

int attention_tc_init() {
  CFM_SMEM_OPT_IN(attention_tc_kernel, kSmemBytes);      // per device: cfm_init runs once per device
  return 0;
}

int attention_tc(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts, const void* v,
                 int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk, const uint8_t* mask, int64_t mask_bs,
                 int64_t mask_rs, const float* key_bias, float scale, int dtype, cudaStream_t st) {
  CFM_CHECK_ARG(key_bias == nullptr, "cfm_attention(tc): key_bias is only supported by the SIMT engine");
  CFM_CHECK_ARG(scale > 0.f, "cfm_attention(tc): scale must be positive");
  CFM_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                  reinterpret_cast<uintptr_t>(out)) & 15) == 0, "cfm_attention(tc): q/k/v/out must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_qkv_tmap(&tmQ, q, B, Tq, H, q_bs, q_ts)) != 0) return rc;
  if ((rc = make_qkv_tmap(&tmK, k, B, Tk, H, k_bs, k_ts)) != 0) return rc;
  if ((rc = make_qkv_tmap(&tmV, v, B, Tk, H, v_bs, v_ts)) != 0) return rc;
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.mask = mask; p.mask_bs = mask_bs; p.mask_rs = mask_rs;
  p.H = H; p.Tq = Tq; p.Tk = Tk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = nullptr;
  if (const char* e = getenv("CFM_B200_ATTN_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  p.mask_aligned8 = (mask != nullptr) && ((reinterpret_cast<uintptr_t>(mask) | (uintptr_t)mask_bs | (uintptr_t)mask_rs) % 8 == 0);
  dim3 grid((Tq + QT - 1) / QT, H, B);
  CFM_CUDA_OK(launch_pdl(attention_tc_kernel, grid, dim3(kThreads), kSmemBytes, st, 1, tmQ, tmK, tmV, p));
  CFM_LAUNCHED_K("attention_tc");
  return 0;
}

}  // namespace cfm
