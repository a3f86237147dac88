// Fused multi-head self-attention + output projection for sm_100a (d = 256 = 4 heads x 64, Tk <= 256):
//     X += W_o . concat_h softmax(mask(q_h k_h^T * scale)) v_h + b_o            [+ the LayerNorm that follows]
// replacing attention.py:84-99 (scores, masked softmax, context, linear_out) + encoder_layer.py:62-63 of the reference:
// one launch instead of the flash-attention kernel + the residual GEMM, and the (tokens x d) context never leaves the SM.
//
// One CTA = 128 query rows of one utterance, all four heads.  Because every key of the utterance fits one 256-column
// score tile there is no online rescaling:
//   S_h = Q_h K_h^T           tcgen05.mma M=128 N=256 K=64          -> TMEM cols [0,256)
//   P_h = softmax             8 warps, thread = (row, half of the keys): masked row max (halves exchanged through shared
//                             memory), ex2 -> bf16 P in the swizzled A-operand layout, partial row sums in registers;
//                             mask bytes -> -inf and masked probabilities -> 0 exactly like attention.py:89-92
//                             (a fully masked row yields 0)
//   O_h = P_h V_h             tcgen05.mma M=128 N=64 K=256 (V tile straight from TMA as MN-major B) -> TMEM cols 256 + 64 h
//   ctx = [O_h / l_h]_h       bf16 into the A-operand layout (over the dead P tile)
//   acc = ctx W_o^T           tcgen05.mma M=128 N=256 K=256 (re-uses the S columns), W_o streamed over the dead K/V stages
//   shared residual/LayerNorm epilogue (resid_epilogue.cuh) through 3-D tensor maps (rows past the utterance are clipped).
// Q/K/V of head h+1 are prefetched (two 80 KB stages) while head h is processed; S_{h+1} is issued before P_h V_h so the
// softmax warps never wait for the tensor pipe longer than one S tile.
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-11: softmax / epilogue
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include "resid_epilogue.cuh"
#include <math_constants.h>
#include <stdlib.h>

namespace cfm {
namespace {

using namespace tc;

constexpr int NH = 4, DK = 64, D = 256, QT = 128, KT = 256;
constexpr int kAtom = 16384;                   // 128 rows x 128 B
constexpr int kQBytes = kAtom, kKBytes = 2 * kAtom, kVBytes = 2 * kAtom;
constexpr int kStage = kQBytes + kKBytes + kVBytes;          // 80 KB
constexpr int kPBytes = 4 * kAtom;                            // 64 KB: P (4 key atoms) / ctx (4 head atoms)
constexpr int kThreads = 384;
constexpr int kSmemBytes = 2 * kStage + kPBytes + 2048 /*row max / row sum exchange*/ + 512;
static_assert(kSmemBytes <= 232448, "smem budget");

struct MhsaParams {
  const uint8_t* mask;
  int64_t mask_bs, mask_rs;
  const float* bo;
  const float* g1; const float* be1;
  const uint8_t* y_row_valid;
  float scale_log2, eps;
  int B, T, ln_mode, mask_aligned8;
  long long* trace;
  unsigned long long* gtrace;   // optional [launch][cta][4] %globaltimer stamps (tools/launch_timeline.py)
  int gslot;
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

#define MTR(i) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 3) p.trace[i] = clock64(); } while (0)
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GTR(i) do { if (p.gtrace && threadIdx.x == 0) \
  p.gtrace[((size_t)p.gslot * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x) * 4 + (i)] = gtimer(); } while (0)

__global__ void __launch_bounds__(kThreads, 1)
mhsa_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV,      // (H*64, T, B) bf16, box 64 x 128 x 1
                  const __grid_constant__ CUtensorMap tmWo,     // (256, 256) bf16, box 64 x 256
                  const __grid_constant__ CUtensorMap tmX,      // X (256, T, B) fp32, box 32 x 128 x 1 (store)
                  const __grid_constant__ CUtensorMap tmR,      // residual load (same tensor)
                  const __grid_constant__ CUtensorMap tmY,      // y out (256, T, B) bf16, box 64 x 128 x 1
                  const MhsaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  GTR(0);
  uint8_t* sStage = smem;                       // 2 x {Q 16 KB, K 32 KB, V 32 KB}; later W_o pieces, staging rings, parameters
  uint8_t* sP = smem + 2 * kStage;              // P tile -> ctx tile
  float* xch_m = reinterpret_cast<float*>(sP + kPBytes);        // [2][128] partial row maxima
  float* xch_l = xch_m + 256;                                   // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch_l + 256);
  uint64_t* kv_full = bars;                     // [2]
  uint64_t* kv_empty = kv_full + 2;             // [2] P_h V_h retired: stage (and P tile) free
  uint64_t* s_full = kv_empty + 2;              // S_h complete
  uint64_t* p_ready = s_full + 1;               // 256 arrivals: P_h written, S_h read
  uint64_t* pv_done = p_ready + 1;              // P_h V_h retired
  uint64_t* ctx_ready = pv_done + 1;            // 256 arrivals
  uint64_t* w_full = ctx_ready + 1;             // [4] W_o pieces
  uint64_t* acc_full = w_full + 4;
  uint64_t* res_bar = acc_full + 1;             // [2][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);
  uint32_t* svis = tmem_slot + 1;               // [8] visibility words of a broadcast (B,1,Tk) mask

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * QT, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmWo);
    prefetch_tmap(&tmX); prefetch_tmap(&tmR); prefetch_tmap(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 256); mbar_init(pv_done, 1); mbar_init(ctx_ready, 256);
    for (int s = 0; s < 4; ++s) mbar_init(w_full + s, 1);
    mbar_init(acc_full, 1);
    for (int s = 0; s < 8; ++s) mbar_init(res_bar + s, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  GTR(1);
  pdl_wait();
  GTR(2);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 0) {
    // ===================== TMA producer =====================
    for (int h = 0; h < NH; ++h) {
      const int s = h & 1;
      if (h >= 2) mbar_wait(kv_empty + s, 0);               // P_{h-2} V_{h-2} retired
      if (elect_one()) {
        uint8_t* st = sStage + s * kStage;
        mbar_expect_tx(kv_full + s, kStage);
        tma_load_3d(st, &tmQ, kv_full + s, h * DK, i0, b);
        tma_load_3d(st + kQBytes, &tmK, kv_full + s, h * DK, 0, b);
        tma_load_3d(st + kQBytes + kAtom, &tmK, kv_full + s, h * DK, 128, b);
        tma_load_3d(st + kQBytes + kKBytes, &tmV, kv_full + s, h * DK, 0, b);
        tma_load_3d(st + kQBytes + kKBytes + kAtom, &tmV, kv_full + s, h * DK, 128, b);
      }
      __syncwarp();
    }
    // W_o pieces [256 out x 64 in] into the K / V slots of the stages as they die
    for (int pc = 0; pc < 4; ++pc) {
      const int s = pc >> 1;
      if ((pc & 1) == 0) mbar_wait(kv_empty + s, 1);        // P_2 V_2 (stage 0) / P_3 V_3 (stage 1) retired
      if (elect_one()) {
        uint8_t* dst = sStage + s * kStage + kQBytes + (pc & 1) * kKBytes;
        mbar_expect_tx(w_full + pc, kKBytes);
        tma_load_2d(dst, &tmWo, w_full + pc, pc * 64, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(QT, KT, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(QT, DK, 1);    // B = V tile, MN-major (dk contiguous)
    constexpr uint32_t idesc_w = umma_idesc_bf16(QT, D, 0);
    const uint32_t p_addr = smem_u32(sP);
    auto issue_s = [&](int h) {
      const int s = h & 1;
      mbar_wait(kv_full + s, h >> 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(sStage + s * kStage);
        const uint64_t dq = umma_desc_sw128(st), dk = umma_desc_sw128(st + kQBytes);
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) umma_bf16(tmem_base, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int h = 0; h < NH; ++h) {
      mbar_wait(p_ready, h & 1);                            // P_h in smem, S_h read out
      tc_fence_after();
      if (h + 1 < NH) issue_s(h + 1);                       // the softmax warps get their next tile first
      if (elect_one()) {
        const uint64_t dv = umma_desc_sw128(smem_u32(sStage + (h & 1) * kStage + kQBytes + kKBytes));
#pragma unroll
        for (int k = 0; k < KT / 16; ++k) {
          // A: P, K-major, four 64-key swizzle atoms, 32 bytes per K step inside an atom
          const uint64_t da = umma_desc_sw128(p_addr + (k >> 2) * kAtom) + 2 * (k & 3);
          // B: V tile rows = keys (128 bytes each): 16 keys per K step = 2048 bytes = +128 in 16-byte units
          umma_bf16(tmem_o + h * DK, da, dv + 128 * k, idesc_o, k != 0);
        }
        umma_commit(kv_empty + (h & 1));
        umma_commit(pv_done);
      }
      __syncwarp();
    }
    // output projection on the normalised context
    mbar_wait(ctx_ready, 0);
    tc_fence_after();
    for (int pc = 0; pc < 4; ++pc) {
      mbar_wait(w_full + pc, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_sw128(p_addr + pc * kAtom);
        const uint64_t db = umma_desc_sw128(smem_u32(sStage + (pc >> 1) * kStage + kQBytes + (pc & 1) * kKBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc_w, (pc | k) != 0);
        if (pc == 3) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== softmax / epilogue: thread = (query row, key half) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - grp * 128;
    const int tid = threadIdx.x - 128;
    const bool elected = (et == 0);
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int i = i0 + r;
    const bool row_ok = i < p.T;
    uint32_t ring_phase = 0;

    // visibility bits of this row for its 4 x 32 keys (the same for every head)
    uint32_t vis[4];
    if (p.mask != nullptr && p.mask_rs == 0) {
      // (B,1,Tk) key-padding mask: the 256 threads fetch one byte each and ballot -> 8 words shared by all rows
      const bool on = (tid < p.T) && (__ldg(p.mask + b * p.mask_bs + tid) != 0);
      const uint32_t w = __ballot_sync(0xffffffffu, on);
      if (lane == 0) svis[tid >> 5] = w;
      named_bar_sync(3, 256);
#pragma unroll
      for (int c = 0; c < 4; ++c) vis[c] = row_ok ? svis[grp * 4 + c] : 0u;
    } else {
      const uint8_t* mrow = (p.mask != nullptr && row_ok) ? p.mask + b * p.mask_bs + i * p.mask_rs : nullptr;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = grp * 128 + c * 32;
        const int nvalid = p.T - jj;
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
    const uint32_t tmem_srow = tmem_base + lane_base + grp * 128;
    // (the head loop is deliberately NOT unrolled: the kernel runs once per layer between other large kernels, so its
    //  code is fetched cold and straight-line copies of the two passes cost more than they save)
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;              // complete row sums per head
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      if (tid == 0) MTR(8 + 4 * h);
      mbar_wait(s_full, h & 1);
      tc_fence_after();
      if (tid == 0) MTR(9 + 4 * h);
      // pass 1: masked maximum of this thread's 128 scores
      float m_part = -CUDART_INF_F;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_srow + c * 32, v);
        tmem_ld_wait();
        if (vis[c] == 0xffffffffu) {
          // four independent chains instead of one 32-deep FMNMX dependency chain
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
          for (int e = 4; e < 32; e += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[e])); m1 = fmaxf(m1, __uint_as_float(v[e + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[e + 2])); m3 = fmaxf(m3, __uint_as_float(v[e + 3]));
          }
          m_part = fmaxf(m_part, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        } else if (vis[c] != 0u) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if ((vis[c] >> e) & 1u) m_part = fmaxf(m_part, __uint_as_float(v[e]));
        }
      }
      xch_m[grp * 128 + r] = m_part;
      named_bar_sync(3, 256);
      const float m_row = fmaxf(m_part, xch_m[(grp ^ 1) * 128 + r]) * p.scale_log2;   // scale > 0 commutes with max
      const float neg_m = (m_row != -CUDART_INF_F) ? -m_row : 0.f;
      if (h > 0) mbar_wait(pv_done, (h - 1) & 1);           // P_{h-1} V_{h-1} has finished reading the P tile
      if (tid == 0) MTR(10 + 4 * h);
      // pass 2: p = 2^(s*c - m); masked keys become -inf first so they come out as exactly 0
      float l = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_srow + c * 32, v);
        tmem_ld_wait();
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
        l += la + lb;
        // this row's 16-byte chunks 4c'..4c'+3 (c' = 4 grp + c) of the 32-chunk P row: atom = c'/2
        const int cc0 = 4 * (grp * 4 + c);
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int cc = cc0 + qq;
          *reinterpret_cast<uint4*>(sP + (cc >> 3) * kAtom + sw_off(r, cc & 7)) =
              make_uint4(pk[4 * qq], pk[4 * qq + 1], pk[4 * qq + 2], pk[4 * qq + 3]);
        }
      }
      xch_l[grp * 128 + r] = l;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      if (tid == 0) MTR(11 + 4 * h);
      named_bar_sync(3, 256);
      const float lt = l + xch_l[(grp ^ 1) * 128 + r];
      l0 = h == 0 ? lt : l0; l1 = h == 1 ? lt : l1; l2 = h == 2 ? lt : l2; l3 = h == 3 ? lt : l3;
    }
    // The residual rows of the epilogue: each warpgroup's staging ring is its stage's [Q | K | V] area.  Both Q slots are
    // dead (S_2 / S_3 have retired), so the first 32-column chunk is fetched now; the other three follow once the
    // projection has finished reading W_o out of the K / V slots.
    uint8_t* ring = sStage + grp * kStage;
    if (elected) {
      mbar_expect_tx(res_bar + grp * 4, QT * 128);
      resid_tma_load(ring, &tmR, res_bar + grp * 4, grp * 128, i0, b);
    }
    // ---- context: O_h / l_h -> bf16 A operand (head h = k atom h); this warpgroup takes 32 of each head's 64 columns
    mbar_wait(pv_done, (NH - 1) & 1);
    tc_fence_after();
    if (tid == 0) MTR(24);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const float lt = h == 0 ? l0 : (h == 1 ? l1 : (h == 2 ? l2 : l3));
      const float inv = lt > 0.f ? 1.f / lt : 0.f;          // fully masked row -> 0 (attention.py:92)
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_base + h * DK + grp * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]) * inv;
        *reinterpret_cast<uint4*>(sP + h * kAtom + sw_off(r, grp * 4 + j)) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(ctx_ready);
    if (tid == 0) MTR(25);
    // ---- projection accumulator -> residual stream (+ LayerNorm); each warpgroup takes 128 columns
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tid == 0) MTR(26);
    // every MMA has retired: the K / V slots (W_o) and the ctx tile (-> parameters, row-sum exchange) are dead
    float* sparam = reinterpret_cast<float*>(sP);
    resid_stage_params<D, 256>(sparam, tid, p.bo, 0, p.ln_mode, p.g1, p.be1, nullptr, nullptr);
    if (elected) {
#pragma unroll
      for (int c = 1; c < 4; ++c) {
        mbar_expect_tx(res_bar + grp * 4 + c, QT * 128);
        resid_tma_load(ring + c * kBufBytes, &tmR, res_bar + grp * 4 + c, (grp * 4 + c) * 32, i0, b);
      }
    }
    ResidParams rp{nullptr, p.y_row_valid, 1.0f, p.eps, p.ln_mode, p.B * p.T};
    ResidOpts ro;
    if (p.trace && blockIdx.x == 0 && blockIdx.y == 3) ro.trace = p.trace + 32;
    resid_ln_epilogue<D, 4, QT, 2>(tmem_base + lane_base, r, i0, 0, elected, 1 + grp, ring, res_bar + grp * 4, ring_phase, sparam,
                                   &tmX, &tmR, &tmY, rp, grp, 3, reinterpret_cast<float2*>(sP + 8192), b, b * p.T + i0, ro);
    if (tid == 0) MTR(27);
    if (elected) bulk_wait_read<0>();   // the stores only have to be done READING shared memory before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  GTR(3);
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int make_qkv_map(CUtensorMap* tm, const void* base, int B, int T, int64_t bs, int64_t ts) {
  const uint64_t dims[3] = {(uint64_t)D, (uint64_t)T, (uint64_t)B};
  const uint64_t str[2] = {(uint64_t)ts * 2, (uint64_t)bs * 2};
  const uint32_t box[3] = {DK, 128, 1};
  return tc::make_tmap_bf16(tm, base, 3, dims, str, box);
}

}  // namespace

bool mhsa_fused_supported(int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts, int B, int H,
                          int Tq, int Tk, int d, int dtype, bool has_key_bias) {
  if (dtype != CFM_BF16 || tc::encode_tiled_fn() == nullptr) return false;
  if (H != NH || d != D || Tq != Tk || Tk > KT || Tq < 64 || has_key_bias) return false;
  if ((q_bs | q_ts | k_bs | k_ts | v_bs | v_ts) % 8 != 0) return false;
  return B >= 1 && B <= 65535;
}

int mhsa_fused(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts, const void* v,
               int64_t v_bs, int64_t v_ts, int B, int T, const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, float scale,
               const void* Wo, const float* bo, float* X, const float* g1, const float* be1, void* Y,
               const uint8_t* y_row_valid, float eps, cudaStream_t st) {
  CFM_CHECK_ARG(scale > 0.f, "cfm_mhsa_out(tc): scale must be positive");
  CFM_SMEM_OPT_IN(mhsa_fused_kernel, kSmemBytes);
  CUtensorMap tmQ, tmK, tmV, tmWo, tmX, tmY;
  int rc;
  if ((rc = make_qkv_map(&tmQ, q, B, T, q_bs, q_ts)) != 0) return rc;
  if ((rc = make_qkv_map(&tmK, k, B, T, k_bs, k_ts)) != 0) return rc;
  if ((rc = make_qkv_map(&tmV, v, B, T, v_bs, v_ts)) != 0) return rc;
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)D};
    const uint64_t str[1] = {(uint64_t)D * 2};
    const uint32_t box[2] = {64, 256};
    if ((rc = tc::make_tmap_bf16(&tmWo, Wo, 2, dims, str, box)) != 0) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)D, (uint64_t)T, (uint64_t)B};
    const uint64_t strx[2] = {(uint64_t)D * 4, (uint64_t)T * D * 4};
    const uint32_t boxx[3] = {32, 128, 1};
    if ((rc = tc::make_tmap_f32(&tmX, X, 3, dims, strx, boxx)) != 0) return rc;
    tmY = tmX;
    if (g1 != nullptr) {
      const uint64_t stry[2] = {(uint64_t)D * 2, (uint64_t)T * D * 2};
      const uint32_t boxy[3] = {64, 128, 1};
      if ((rc = tc::make_tmap_bf16(&tmY, Y, 3, dims, stry, boxy)) != 0) return rc;
    }
  }
  MhsaParams p;
  p.mask = mask; p.mask_bs = mask_bs; p.mask_rs = mask_rs;
  p.bo = bo; p.g1 = g1; p.be1 = be1; p.y_row_valid = y_row_valid;
  p.scale_log2 = scale * 1.4426950408889634f; p.eps = eps;
  p.B = B; p.T = T; p.ln_mode = g1 ? 1 : 0;
  p.mask_aligned8 = (mask != nullptr) && ((reinterpret_cast<uintptr_t>(mask) | (uintptr_t)mask_bs | (uintptr_t)mask_rs) % 8 == 0);
  p.trace = nullptr; p.gtrace = nullptr; p.gslot = 0;
  if (const char* e = getenv("CFM_B200_MHSA_GTRACE_PTR")) {
    static int slot = 0;
    p.gtrace = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
    p.gslot = slot++ % 64;
  }
  if (const char* e = getenv("CFM_B200_MHSA_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  dim3 grid((T + QT - 1) / QT, B);
  CFM_CUDA_OK(launch_pdl(mhsa_fused_kernel, grid, dim3(kThreads), kSmemBytes, st, 1, tmQ, tmK, tmV, tmWo, tmX, tmX, tmY, p));
  CFM_LAUNCHED_K("mhsa_fused");
  return 0;
}

}  // namespace cfm
