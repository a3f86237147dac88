// tcgen05 flash attention for LONG sequences on sm_100a (head dim 64, bf16 operands, fp32 softmax statistics): the
// two-query-tile "ping-pong" organisation.  attention_tc.cu runs S = QK^T -> softmax -> PV strictly in series inside a
// CTA and relies on two CTAs per SM for overlap; here ONE CTA per SM owns TWO 128-row query tiles of one (batch, head)
// and keeps the tensor pipe, the TMA engine and the MUFU busy at the same time:
//   warp 8   TMA producer: Q0, Q1 once; K and V tiles of 128 keys through a 2-stage ring
//   warp 9   MMA issuer:   S_q = Q_q K_j^T (M=128 N=128 K=64) into TMEM S_q; as soon as softmax group q has turned S_q(j)
//                          into P_q(j) it issues O_q = P_q V_j (N=64, V straight from TMA as an MN-major operand) and, right
//                          behind it, S_q(j+1) -- so the next score tile is ready before group q comes back for it
//   warp 10  TMEM allocator (S0, S1: 2 x 128 columns; O0, O1: 2 x 64 columns)
//   warps 0-3 / 4-7   softmax groups 0 / 1, thread = query row: masked row max, ex2, P as bf16 into the swizzled A-operand
//                          layout, running (max, sum) and the 64-wide output row in registers with the online rescale
//                          (mask semantics of attention.py:89-92: masked -> -inf, masked probabilities 0, fully masked row 0)
// While group 0 is in its MUFU-bound ex2 pass, group 1's S / PV MMAs run, and vice versa.  Used for Tq > 128 (C4: T = 1498,
// C3: T = 498); shorter query ranges stay on attention_tc.cu.  No T x T tensor is materialised (attention.py:84-96).
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace cfm {
namespace {

using namespace tc;

constexpr int QT = 128, KT = 128, DK = 64, NQ = 2;
// SPLIT = threads per query row in the softmax groups.  1: 4 warps per group (one per scheduler); 2: 8 warps per group,
// thread = (row, key half / output-column half): four softmax warps per scheduler instead of two hide the shared-memory,
// TMEM and MUFU latencies of the passes (a single warp's ex2 pass is latency bound: 2.0 k cycles alone vs a 1.0 k MUFU
// bound), the halves exchange their row maxima through shared memory once per tile.
template <int SPLIT> struct PPCfg {
  static constexpr int kSoftWarps = 8 * SPLIT;
  static constexpr int kTmaWarp = kSoftWarps, kMmaWarp = kSoftWarps + 1, kAllocWarp = kSoftWarps + 2;
  static constexpr int kThreads = (kSoftWarps + 3) * 32;
};
constexpr int kTile = 128 * DK * 2;            // 16 KB: one Q, K or V tile
constexpr int kPBytes = QT * KT * 2;           // 32 KB per query tile
constexpr int kSmemBytes = NQ * kTile + 2 * 2 * kTile + NQ * kPBytes + 1024 + 4096;   // + row-maximum / row-sum exchange (SPLIT = 2)
constexpr int kTmemCols = 512;                 // 2 x 128 (S) + 2 x 64 (O) = 384 -> next power of two

struct AttnPPParams {
  __nv_bfloat16* out;
  const uint8_t* mask;
  int64_t mask_bs, mask_rs;
  int H, Tq, Tk;
  float scale_log2;
  int mask_aligned8;
  long long* trace;     // optional clock64 timeline of CTA (0,0,0) (tools/attn_pp_trace.py); nullptr in production
};

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

// PTMEM (with SPLIT = 2): P is written into tensor memory (tcgen05.st, 64 columns per query tile: TMEM holds S 2 x 128,
// O 2 x 64, P 2 x 64 = 512 columns) and P V takes its A operand from there: no shared-memory P tile, no
// fence.proxy.async after the stores (measured ~500 cycles per tile on the thread that issues it).
template <int SPLIT, bool PTMEM>
__global__ void __launch_bounds__(PPCfg<SPLIT>::kThreads, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnPPParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;                                   // [2] Q tiles
  uint8_t* sKV = smem + NQ * kTile;                     // [2 stages] {K tile, V tile}
  uint8_t* sP = sKV + 2 * 2 * kTile;                    // [2] P tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NQ * kPBytes);
  uint64_t* q_full = bars;                              // 1
  uint64_t* kv_full = bars + 1;                         // [2]
  uint64_t* kv_empty = bars + 3;                        // [2]
  uint64_t* s_full = bars + 5;                          // [2] per query tile
  uint64_t* p_ready = bars + 7;                         // [2] 128 arrivals
  uint64_t* o_full = bars + 9;                          // [2]
  uint64_t* s_free = bars + 11;                         // [2] SPLIT = 2: S_q read out of TMEM (before P_q is complete)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  uint32_t* svis = tmem_slot + 1;                       // [2 groups][2 parities][4] visibility words of a (B,1,Tk) mask
  float* sxch = reinterpret_cast<float*>(sP + NQ * kPBytes + 1024);   // [2 groups][2 parities][2 halves][128] (SPLIT = 2)
  using PC = PPCfg<SPLIT>;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * (NQ * QT);
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Tk + KT - 1) / KT;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define PTR(slot) do { if (tr && lane == 0) p.trace[(slot)] = clock64(); } while (0)

  if (warp == PC::kTmaWarp && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1);
      mbar_init(s_full + s, 1); mbar_init(p_ready + s, 128 * SPLIT); mbar_init(o_full + s, 1);
      mbar_init(s_free + s, 128 * SPLIT);
    }
    fence_barrier_init();
  }
  if (warp == PC::kAllocWarp) tmem_alloc<kTmemCols>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == PC::kTmaWarp) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(q_full, NQ * kTile);
      tma_load_3d(sQ, &tmQ, q_full, h * DK, i0, b);
      tma_load_3d(sQ + kTile, &tmQ, q_full, h * DK, i0 + QT, b);        // rows past Tq: zero filled
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      if (j >= 2) mbar_wait(kv_empty + s, ((j >> 1) - 1) & 1);          // the MMAs of tile j-2 have retired
      if (elect_one()) {
        uint8_t* st = sKV + s * 2 * kTile;
        mbar_expect_tx(kv_full + s, 2 * kTile);
        tma_load_3d(st, &tmK, kv_full + s, h * DK, j * KT, b);
        tma_load_3d(st + kTile, &tmV, kv_full + s, h * DK, j * KT, b);
      }
      __syncwarp();
    }
  } else if (warp == PC::kMmaWarp) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KT, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(QT, DK, 1);           // B = V tile, MN-major (dk contiguous)
    auto issue_s = [&](int q, int j) {
      if (elect_one()) {
        const uint64_t dq = umma_desc_sw128(smem_u32(sQ + q * kTile));
        const uint64_t dk = umma_desc_sw128(smem_u32(sKV + (j & 1) * 2 * kTile));
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) umma_bf16(tmem_base + q * KT, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full + q);
      }
      __syncwarp();
    };
    PTR(0);
    mbar_wait(q_full, 0);
    mbar_wait(kv_full, 0);
    tc_fence_after();
    PTR(1);
    issue_s(0, 0);
    issue_s(1, 0);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
#pragma unroll 1
      for (int q = 0; q < NQ; ++q) {
        if (SPLIT == 2 && j + 1 < n_kv) {
          // the next score tile goes into the pipe as soon as group q has READ S_q(j) (it is still busy with the ex2 of
          // its last chunk and the P stores): S_q(j+1) is complete when the group comes back for it
          mbar_wait(s_free + q, j & 1);
          if (q == 0) mbar_wait(kv_full + ((j + 1) & 1), ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(q, j + 1);
        }
        mbar_wait(p_ready + q, j & 1);          // P_q(j) in smem, S_q(j) read out, O_q(j-1) drained
        tc_fence_after();
        if (j < 6) PTR(16 + (j * 2 + q) * 2);
        if (elect_one()) {
          const uint64_t dv = umma_desc_sw128(smem_u32(sKV + s * 2 * kTile + kTile));
          const uint32_t pa = smem_u32(sP + q * kPBytes);
#pragma unroll
          for (int k = 0; k < KT / 16; ++k) {
            // B: 16 keys per K step = 2048 bytes = +128 units
            if constexpr (PTMEM) {              // A: P in tensor memory, 8 columns (16 packed bf16) per K step
              umma_bf16_ts(tmem_base + NQ * KT + q * DK, tmem_base + NQ * KT + NQ * DK + q * (KT / 2) + k * 8, dv + 128 * k,
                           idesc_o, k != 0);
            } else {                            // A: P in shared memory, K-major, two 64-key swizzle atoms (16 KB each)
              const uint64_t da = umma_desc_sw128(pa + (k >> 2) * (kPBytes / 2)) + 2 * (k & 3);
              umma_bf16(tmem_base + NQ * KT + q * DK, da, dv + 128 * k, idesc_o, k != 0);
            }
          }
          umma_commit(o_full + q);
          if (q == NQ - 1) umma_commit(kv_empty + s);   // every MMA that reads stage s was issued before this commit
        }
        __syncwarp();
        if (SPLIT == 1 && j + 1 < n_kv) {
          if (q == 0) { mbar_wait(kv_full + ((j + 1) & 1), ((j + 1) >> 1) & 1); tc_fence_after(); }
          issue_s(q, j + 1);                    // S_q is free: group q has read S_q(j) (p_ready above)
        }
        if (j < 6) PTR(16 + (j * 2 + q) * 2 + 1);
      }
    }
  } else if (SPLIT == 2 && warp < PC::kSoftWarps) {
    // ===================== softmax groups, thread = (query row, key half) =====================
    const int g = warp >> 3;                    // group = query tile
    const int wq = warp & 3;                    // TMEM lane quadrant (= warp % 4)
    const int hf = (warp >> 2) & 1;             // key half of the score tile / column half of the output row
    const int r = wq * 32 + lane;
    const int i = i0 + g * QT + r;
    const bool row_ok = i < p.Tq;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * KT + lane_base + hf * 64;
    const uint32_t tmem_o = tmem_base + NQ * KT + g * DK + lane_base + hf * 32;
    uint8_t* sPg = sP + g * kPBytes + hf * (kPBytes / 2);     // K atom `hf` of the P tile
    const uint32_t tmem_p = tmem_base + NQ * KT + NQ * DK + g * (KT / 2) + lane_base + hf * 32;   // PTMEM: 32 packed columns
    const uint8_t* mrow = (p.mask != nullptr && row_ok) ? p.mask + b * p.mask_bs + i * p.mask_rs : nullptr;
    float* xch = sxch + g * 512;
    uint32_t* sv = svis + g * 8;
    float m_run = -CUDART_INF_F, l_run = 0.f;   // l_run: this thread's key half only (summed at the end)
    float o[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = 0.f;
    const bool bcast_mask = p.mask != nullptr && p.mask_rs == 0;
    // (B,1,Tk) key-padding mask: the hf = 0 warps fetch one byte per thread and ballot it into 4 words per tile, one tile
    // ahead; the words of tile j+1 are published by the row-maximum barrier of tile j
    uint8_t mbyte = (bcast_mask && hf == 0 && r < p.Tk) ? __ldg(p.mask + b * p.mask_bs + r) : (uint8_t)0;
    auto publish_vis = [&](int j) {             // visibility words of tile j -> sv[(j & 1) * 4 + ...], next byte prefetched
      if (bcast_mask && hf == 0) {
        const int jj = j * KT + r;
        const bool on = (jj < p.Tk) && (mbyte != 0);
        const int jn = jj + KT;
        mbyte = (jn < p.Tk) ? __ldg(p.mask + b * p.mask_bs + jn) : (uint8_t)0;
        const uint32_t w = __ballot_sync(0xffffffffu, on);
        if (lane == 0) sv[(j & 1) * 4 + wq] = w;
      }
    };
    publish_vis(0);
    named_bar_sync(1 + g, 256);
    float alpha_prev = 1.f;
#pragma unroll 1
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int j0 = j * KT + hf * 64;          // first key of this thread's half
      uint32_t vis[2];
      if (bcast_mask) {
#pragma unroll
        for (int c = 0; c < 2; ++c) vis[c] = row_ok ? sv[(j & 1) * 4 + 2 * hf + c] : 0u;
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
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
      if (j + 1 < n_kv) publish_vis(j + 1);
      if (wq == 0 && hf == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 0);
      mbar_wait(s_full + g, ph);
      tc_fence_after();
      if (wq == 0 && hf == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 1);
      // ---- pass 1: masked maximum of this thread's 64 scores, halves exchanged through shared memory
      float m_tile = -CUDART_INF_F;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tmem_s + half * 32, v);
        tmem_ld_wait();
        const uint32_t vm = vis[half];
        if (vm == 0xffffffffu) {
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
          for (int e = 4; e < 32; e += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[e])); m1 = fmaxf(m1, __uint_as_float(v[e + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[e + 2])); m3 = fmaxf(m3, __uint_as_float(v[e + 3]));
          }
          m_tile = fmaxf(m_tile, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else if (vm != 0u) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((vm >> e) & 1u) m_tile = fmaxf(m_tile, __uint_as_float(v[e]));
        }
      }
      xch[(j & 1) * 256 + hf * 128 + r] = m_tile;
      named_bar_sync(1 + g, 256);               // (also publishes the visibility words of tile j+1)
      m_tile = fmaxf(m_tile, xch[(j & 1) * 256 + (hf ^ 1) * 128 + r]) * p.scale_log2;
      if (wq == 0 && hf == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 2);
      const float m_new = fmaxf(m_run, m_tile);
      const bool any = m_new != -CUDART_INF_F;
      const float alpha = any ? exp2f(m_run - m_new) : 1.f;
      const float neg_m = any ? -m_new : 0.f;
      // O_tile(j-1) -> registers with ITS rescale factor (this thread's 32 output columns)
      if (j > 0) {
        mbar_wait(o_full + g, (j - 1) & 1);
        tc_fence_after();
        if (wq == 0 && hf == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 4);
        uint32_t va[32];
        tmem_ld32(tmem_o, va);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], alpha_prev, __uint_as_float(va[e]));
        if (wq == 0 && hf == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 5);
      }
      // group 1 starts its first ex2 pass after group 0's (see the SPLIT = 1 path)
      if (g == 1 && j == 0) asm volatile("bar.sync 3, 512;" ::: "memory");
      // ---- pass 2: p = 2^(s * scale - m), bf16 P into K atom `hf`, partial row sum
      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];                         // (one chunk at a time: 96 registers per thread at 608 threads)
        tmem_ld32(tmem_s + c * 32, v);
        tmem_ld_wait();
        if (c == 1) { tc_fence_before(); mbar_arrive(s_free + g); }     // S_g(j) is out of TMEM: S_g(j+1) may overwrite it
        if (vis[c] != 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((vis[c] >> e) & 1u)) v[e] = 0xff800000u;  // -inf
        }
        uint32_t pk[16];
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float p0 = ex2_fast(fmaf(__uint_as_float(v[e]), p.scale_log2, neg_m));
          const float p1 = ex2_fast(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, neg_m));
          const float p2 = ex2_fast(fmaf(__uint_as_float(v[e + 2]), p.scale_log2, neg_m));
          const float p3 = ex2_fast(fmaf(__uint_as_float(v[e + 3]), p.scale_log2, neg_m));
          la += p0 + p1;
          lb += p2 + p3;
          pk[e >> 1] = pack_bf16x2(p0, p1);
          pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
        }
        l_tile += la + lb;
        if constexpr (PTMEM) {
          tmem_st16(tmem_p + c * 16, pk);
        } else {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int cc = 4 * c + qq;            // 16-byte chunk inside this half's 128-byte row
            uint8_t* dst = sPg + r * 128 + ((cc ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
          }
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      if constexpr (PTMEM) tmem_st_wait(); else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready + g);
      if (g == 0 && j == 0) asm volatile("bar.arrive 3, 512;" ::: "memory");
      if (wq == 0 && hf == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 3);
      alpha_prev = alpha;
    }
    // the last tile's P V product, and the row sum of the other key half
    float* lx = xch + (n_kv & 1) * 256;          // the parity slot the last tile's row-maximum exchange did NOT use
    lx[hf * 128 + r] = l_run;
    mbar_wait(o_full + g, (n_kv - 1) & 1);
    tc_fence_after();
    {
      uint32_t va[32];
      tmem_ld32(tmem_o, va);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], alpha_prev, __uint_as_float(va[e]));
    }
    tc_fence_before();
    named_bar_sync(1 + g, 256);
    l_run += lx[(hf ^ 1) * 128 + r];
    if (row_ok) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.Tq + i) * p.H * DK + h * DK + hf * 32);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dst[c] = make_uint4(pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                            pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
  } else if (SPLIT == 1 && warp < 8) {
    // ===================== softmax groups: thread = query row =====================
    const int g = warp >> 2;                    // group = query tile
    const int wq = warp & 3;                    // TMEM lane quadrant
    const int r = wq * 32 + lane;
    const int i = i0 + g * QT + r;
    const bool row_ok = i < p.Tq;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tmem_s = tmem_base + g * KT + lane_base;
    const uint32_t tmem_o = tmem_base + NQ * KT + g * DK + lane_base;
    uint8_t* sPg = sP + g * kPBytes;
    const uint8_t* mrow = (p.mask != nullptr && row_ok) ? p.mask + b * p.mask_bs + i * p.mask_rs : nullptr;
    float m_run = -CUDART_INF_F, l_run = 0.f;
    float o[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) o[c] = 0.f;

    // Started together the two groups stay in phase: they fight for the MUFU during their ex2 passes and idle together
    // while the P V products are in flight (clock trace: 4.35 k cycles per tile and group).  Group 1 therefore starts one
    // ex2 pass late (see below): 4.0 k.  Measured dead ends (tools/attn_pp_trace.py): a strict token
    // that lets only one group into the ex2 pass at a time (a pass takes 2.0 k cycles alone, 2.3 k for both groups
    // together: the pass is latency / issue bound per warp, not MUFU bound), and a cubic-polynomial ex2 on the FMA pipe
    // for a quarter or half of the elements (more instructions per warp: 4.9 k).
    const bool bcast_mask = p.mask != nullptr && p.mask_rs == 0;
    // (B,1,Tk) key-padding mask: each thread fetches one byte per tile, ONE TILE AHEAD (the dependent global load used to
    // sit at the top of every iteration: ~600 exposed cycles per tile)
    uint8_t mbyte = (bcast_mask && r < p.Tk) ? __ldg(p.mask + b * p.mask_bs + r) : (uint8_t)0;
    float alpha_prev = 1.f;
#pragma unroll 1
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t ph = j & 1;
      const int j0 = j * KT;
      uint32_t vis[4];
      if (bcast_mask) {
        const int jj = j0 + r;
        const bool on = (jj < p.Tk) && (mbyte != 0);
        {
          const int jn = jj + KT;
          mbyte = (j + 1 < n_kv && jn < p.Tk) ? __ldg(p.mask + b * p.mask_bs + jn) : (uint8_t)0;
        }
        const uint32_t w = __ballot_sync(0xffffffffu, on);
        uint32_t* sv = svis + (g * 2 + (j & 1)) * 4;
        if (lane == 0) sv[wq] = w;
        named_bar_sync(1 + g, 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) vis[c] = row_ok ? sv[c] : 0u;
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
      if (wq == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 0);
      mbar_wait(s_full + g, ph);
      tc_fence_after();
      if (wq == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 1);
      // Only two warps share a scheduler here, so the ~250-cycle latency of every tcgen05.ld is exposed unless the loads
      // are batched / software-pipelined in registers: pass 1 fetches the row in two 64-column halves (2 exposed
      // latencies instead of 4), pass 2 fetches chunk c+1 before it computes chunk c (1 exposed latency instead of 4).
      float m_tile = -CUDART_INF_F;
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t va[32], vb[32];
        tmem_ld32(tmem_s + (2 * hc) * 32, va);
        tmem_ld32(tmem_s + (2 * hc + 1) * 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t vm = vis[2 * hc + half];
          const uint32_t (&v)[32] = half == 0 ? va : vb;
          if (vm == 0xffffffffu) {
            // four independent chains: a single running max is a 32-deep dependency chain that two warps per
            // scheduler cannot hide
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
            for (int e = 4; e < 32; e += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[e])); m1 = fmaxf(m1, __uint_as_float(v[e + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[e + 2])); m3 = fmaxf(m3, __uint_as_float(v[e + 3]));
            }
            m_tile = fmaxf(m_tile, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else if (vm != 0u) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if ((vm >> e) & 1u) m_tile = fmaxf(m_tile, __uint_as_float(v[e]));
          }
        }
      }
      m_tile *= p.scale_log2;
      if (wq == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 2);
      const float m_new = fmaxf(m_run, m_tile);
      const bool any = m_new != -CUDART_INF_F;
      const float alpha = any ? exp2f(m_run - m_new) : 1.f;
      const float neg_m = any ? -m_new : 0.f;
      // O_tile(j-1) = P(j-1) V(j-1) -> registers with ITS rescale factor, now (it was issued when P(j-1) was published and has
      // had the whole max pass of this tile to complete); must be consumed before P(j) is published (P V(j) overwrites it)
      if (j > 0) {
        mbar_wait(o_full + g, (j - 1) & 1);
        tc_fence_after();
        if (wq == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 4);
        uint32_t va[32], vb[32];
        tmem_ld32(tmem_o, va);
        tmem_ld32(tmem_o + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], alpha_prev, __uint_as_float(va[e]));
#pragma unroll
        for (int e = 0; e < 32; ++e) o[32 + e] = fmaf(o[32 + e], alpha_prev, __uint_as_float(vb[e]));
        if (wq == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 5);
      }
      float l_tile = 0.f;
      uint32_t vbuf[2][32];
      tmem_ld32(tmem_s, vbuf[0]);
      tmem_ld_wait();
      // group 1 starts its first ex2 pass after group 0's: from then on one group's MUFU-bound pass overlaps the other's
      // S wait / row maximum / O fold instead of both fighting for the MUFU and idling together
      if (g == 1 && j == 0) asm volatile("bar.sync 3, 256;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t (&v)[32] = vbuf[c & 1];
        if (c + 1 < 4) tmem_ld32(tmem_s + (c + 1) * 32, vbuf[(c + 1) & 1]);     // in flight while chunk c is computed
        if (vis[c] != 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (!((vis[c] >> e) & 1u)) v[e] = 0xff800000u;  // -inf
        }
        uint32_t pk[16];
        float la = 0.f, lb = 0.f;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          const float p0 = ex2_fast(fmaf(__uint_as_float(v[e]), p.scale_log2, neg_m));
          const float p1 = ex2_fast(fmaf(__uint_as_float(v[e + 1]), p.scale_log2, neg_m));
          const float p2 = ex2_fast(fmaf(__uint_as_float(v[e + 2]), p.scale_log2, neg_m));
          const float p3 = ex2_fast(fmaf(__uint_as_float(v[e + 3]), p.scale_log2, neg_m));
          la += p0 + p1;
          lb += p2 + p3;
          pk[e >> 1] = pack_bf16x2(p0, p1);
          pk[(e >> 1) + 1] = pack_bf16x2(p2, p3);
        }
        l_tile += la + lb;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int cc = 4 * c + qq;
          uint8_t* dst = sPg + (cc >> 3) * (kPBytes / 2) + r * 128 + (((cc & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
        }
        if (c + 1 < 4) tmem_ld_wait();
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready + g);
      if (g == 0 && j == 0) asm volatile("bar.arrive 3, 256;" ::: "memory");
      if (wq == 0 && j < 6) PTR(48 + (g * 6 + j) * 6 + 3);
      alpha_prev = alpha;
    }
    // the last tile's P V product
    mbar_wait(o_full + g, (n_kv - 1) & 1);
    tc_fence_after();
    {
      uint32_t va[32], vb[32];
      tmem_ld32(tmem_o, va);
      tmem_ld32(tmem_o + 32, vb);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], alpha_prev, __uint_as_float(va[e]));
#pragma unroll
      for (int e = 0; e < 32; ++e) o[32 + e] = fmaf(o[32 + e], alpha_prev, __uint_as_float(vb[e]));
    }
    tc_fence_before();
    if (row_ok) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.Tq + i) * p.H * DK + h * DK);
#pragma unroll
      for (int c = 0; c < DK / 8; ++c)
        dst[c] = make_uint4(pack_bf16x2(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16x2(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                            pack_bf16x2(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16x2(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PC::kAllocWarp) tmem_dealloc<kTmemCols>(tmem_base);
}

int make_tmap3(CUtensorMap* tm, const void* base, int B, int T, int H, int64_t bs, int64_t ts) {
  const uint64_t dims[3] = {(uint64_t)H * DK, (uint64_t)T, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)ts * 2, (uint64_t)bs * 2};
  const uint32_t box[3] = {DK, 128, 1};
  return tc::make_tmap_bf16(tm, base, 3, dims, str, box);
}

}  // namespace

int attention_pp(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts, const void* v,
                 int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk, const uint8_t* mask, int64_t mask_bs,
                 int64_t mask_rs, float scale, cudaStream_t st) {
  CFM_CHECK_ARG(scale > 0.f, "cfm_attention(pp): scale must be positive");
  CFM_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                  reinterpret_cast<uintptr_t>(out)) & 15) == 0, "cfm_attention(pp): q/k/v/out must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_tmap3(&tmQ, q, B, Tq, H, q_bs, q_ts)) != 0) return rc;
  if ((rc = make_tmap3(&tmK, k, B, Tk, H, k_bs, k_ts)) != 0) return rc;
  if ((rc = make_tmap3(&tmV, v, B, Tk, H, v_bs, v_ts)) != 0) return rc;
  AttnPPParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.mask = mask; p.mask_bs = mask_bs; p.mask_rs = mask_rs;
  p.H = H; p.Tq = Tq; p.Tk = Tk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.trace = nullptr;
  if (const char* e = getenv("CFM_B200_ATTN_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  p.mask_aligned8 = (mask != nullptr) && ((reinterpret_cast<uintptr_t>(mask) | (uintptr_t)mask_bs | (uintptr_t)mask_rs) % 8 == 0);
  dim3 grid((Tq + NQ * QT - 1) / (NQ * QT), H, B);
  static const bool split1 = env_is("CFM_B200_ATTN_PP_SPLIT", "1"), psmem = env_is("CFM_B200_ATTN_PP_PTMEM", "0");
  if (split1) {
    CFM_SMEM_OPT_IN((attention_pp_kernel<1, false>), kSmemBytes);
    CFM_CUDA_OK(launch_pdl(attention_pp_kernel<1, false>, grid, dim3(PPCfg<1>::kThreads), kSmemBytes, st, 1, tmQ, tmK, tmV, p));
  } else if (psmem) {
    CFM_SMEM_OPT_IN((attention_pp_kernel<2, false>), kSmemBytes);
    CFM_CUDA_OK(launch_pdl(attention_pp_kernel<2, false>, grid, dim3(PPCfg<2>::kThreads), kSmemBytes, st, 1, tmQ, tmK, tmV, p));
  } else {
    CFM_SMEM_OPT_IN((attention_pp_kernel<2, true>), kSmemBytes);
    CFM_CUDA_OK(launch_pdl(attention_pp_kernel<2, true>, grid, dim3(PPCfg<2>::kThreads), kSmemBytes, st, 1, tmQ, tmK, tmV, p));
  }
  CFM_LAUNCHED_K("attention_pp");
  return 0;
}

}  // namespace cfm
