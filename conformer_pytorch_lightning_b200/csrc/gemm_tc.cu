// tcgen05 GEMM engine for sm_100a:  C = epilogue(A W^T + bias), A (M,K) bf16, W (N,K) bf16 (both K-major),
// fp32 accumulation in TMEM.
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0 (1 thread)  TMA producer: A tile 128x64 and W tile BNx64 (128-byte swizzle) into a
//                      STAGES-deep shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1 (1 thread)  MMA issuer: 4 x tcgen05.mma (M=128, N=BN, K=16) per stage into one of two TMEM
//                      accumulator buffers; tcgen05.commit releases the smem slot / publishes the tile
//   warp 2             TMEM allocator (2 x BN fp32 columns)
//   warps 4-11         epilogue, thread = output row x column half (tcgen05.ld 32x32b).  All global traffic of the epilogue is
//                      bulk-asynchronous: results are written into 128-byte-swizzled staging tiles in shared
//                      memory and leave through TMA stores; the fp32 residual tile arrives through TMA loads that
//                      are prefetched while the main loop is still running.
// Epilogues:
//   BIAS / BIAS_SILU / BIAS_GLU  -> bf16 activations                       (feedforward.py:17-18, attention.py:62-64,
//                                                                           convolution.py:41-42)
//   RESIDUAL                     -> X = R + alpha * rowmask(acc + bias)     (encoder_layer.py:58,62,66,69)
//   RESIDUAL + LN  (N == tile)   -> additionally y = LN(X) (bf16, optional row mask), or X = LN1(.), y = LN2(X):
//                                   the LayerNorms of encoder_layer.py:59,63,67,70 / 56 fused into the producing GEMM;
//                                   row statistics are thread-local because a thread owns a whole output row, and the
//                                   pre-norm row is parked in the tile's own TMEM accumulator between the passes.
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include "resid_epilogue.cuh"
#include <math_constants.h>
#include <stdlib.h>

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

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  CFM_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  CFM_CHECK_ARG(rank >= 2 && rank <= 5, "tensor map rank %d unsupported", rank);
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFM_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
                (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return 0;
}
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int BM = 128;       // rows per tile  = UMMA M
constexpr int BK = 64;        // K per stage    = one 128-byte swizzle atom of bf16
constexpr int UK = 16;        // K per tcgen05.mma (bf16)
constexpr int kMaxSmem = 232448;

// PAIR: the tile is 256 rows x BN columns on the two SMs of a 2-CTA cluster (tcgen05 cta_group::2): every CTA stages its own
// 128 A rows and only HALF of the W rows, so the shared-memory fill and the B-operand reads per SM are halved (the
// single-CTA kernel reads 96 B/clk of operands and fills 96 B/clk by TMA against the 128 B/clk one SM's shared memory
// serves) and the smaller stage buys a deeper ring.
template <int BN, bool RESID, bool PAIR = false, int NG = 2> struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue warps: every epilogue is latency bound with one warp per scheduler, so two warpgroups split the tile's
  // columns (2 warps / SMSP); the residual/LN epilogue exchanges its partial row sums through shared memory
  // NG = 4 for the plain activation epilogues of 256-wide tiles: bias + SiLU of a 128 x 256 tile takes two warpgroups 4.1 k
  // cycles (tanh.approx on 2 warps per scheduler), as long as the tile's MMAs at K = 512; four warpgroups own one 64-column
  // sub-tile each
  static constexpr int kGroups = NG;
  static constexpr int kEpiThreads = 128 * NG;
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kBufs = 4;                                  // staging ring (2 per warpgroup)
  static constexpr int kParamFloats = RESID ? 5 * BN + 1024 : BN;  // bias (+ LN gammas/betas + row-sum exchange)
  static constexpr int kFixed = kBufs * kBufBytes + kParamFloats * 4 + 256 /*barriers*/ + 1024 /*align slack*/;
  static constexpr int kStages = (kMaxSmem - kFixed) / kStageBytes > (PAIR ? 8 : 6) ? (PAIR ? 8 : 6) : (kMaxSmem - kFixed) / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(kSmemBytes <= kMaxSmem, "shared memory budget exceeded");
};

struct GemmParams {
  const float* bias;
  const uint8_t* row_valid;     // RESIDUAL: rows whose GEMM result is forced to 0 (pad mask)
  const uint8_t* y_row_valid;   // LN modes: rows of y forced to 0
  const float* g1; const float* b1; const float* g2; const float* b2;
  float alpha, eps;
  int M, N, K;                  // N = number of OUTPUT columns (GLU: W has 2N rows)
  int ln_mode;                  // 0 none, 1 y = LN1(X), 2 X = LN1(.), y = LN2(X)
  int no_resid;                 // RESIDUAL epilogue without a residual operand: X = alpha * rowmask(acc + bias) (fp32 output)
  unsigned long long* keys;     // ARGMAX epilogue: per-row packed (ordered logit, ~column), combined with atomicMax
  int n_valid;                  // ARGMAX: columns >= n_valid are padding (W rows zero-filled by TMA)
  int pair;                     // host-side: launch the cta_group::2 kernel (W tensor map holds half-tile boxes)
  long long* trace;             // optional cycle accounting of CTA 0 (tools/gemm_trace.py); nullptr in production
};
constexpr int EPI_ARGMAX = 100; // internal epilogue of cfm_ctc_argmax: no C tile at all



template <int BN, int EPI> constexpr int epi_groups() {
  return (BN == 256 && (EPI == CFM_EPI_BIAS || EPI == CFM_EPI_BIAS_SILU || EPI == CFM_EPI_BIAS_RELU)) ? 4 : 2;
}

template <int BN, int EPI, bool PAIR>
__global__ void __launch_bounds__((Cfg<BN, EPI == CFM_EPI_RESIDUAL, PAIR, epi_groups<BN, EPI>()>::kThreads), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmC,   // F1: bf16 output; RESIDUAL: fp32 X store
               const __grid_constant__ CUtensorMap tmR,   // RESIDUAL: fp32 residual load
               const __grid_constant__ CUtensorMap tmY,   // LN modes: bf16 y store
               const GemmParams p) {
  constexpr bool GLU = (EPI == CFM_EPI_BIAS_GLU);
  constexpr bool RESID = (EPI == CFM_EPI_RESIDUAL);
  using C = Cfg<BN, RESID, PAIR, epi_groups<BN, EPI>()>;
  constexpr int NG = C::kGroups;
  constexpr int OUT_BN = GLU ? BN / 2 : BN;      // output columns per tile
  // 1024-byte alignment is what SWIZZLE_128B tiles need; keeping `smem` a plain shared-space array (no integer
  // round-up) lets ptxas emit LDS/STS instead of generic LD.E/ST.E for every epilogue access
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + C::kStages * C::kStageBytes;                       // kBufs x 16 KB, 1024-aligned
  float* sparam = reinterpret_cast<float*>(ring + C::kBufs * kBufBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sparam + C::kParamFloats);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tfull_bar = empty_bar + C::kStages;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint64_t* res_bar = tempty_bar + 2;              // [kBufs] residual chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + C::kBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TM = PAIR ? 2 * BM : BM;            // rows per scheduled tile (pair: 128 per CTA)
  const int m_tiles = (p.M + TM - 1) / TM;
  const int n_tiles = p.N / OUT_BN;
  const int total = m_tiles * n_tiles;
  const int kb_count = p.K / BK;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int m_off = (int)crank * BM;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmC);
    if constexpr (RESID) { prefetch_tmap(&tmR); if (p.ln_mode) prefetch_tmap(&tmY); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, (PAIR ? 2 : 1) * C::kEpiThreads); }
    for (int s = 0; s < C::kBufs; ++s) mbar_init(res_bar + s, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_2sm<C::kTmemCols>(tmem_slot);
    else tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();           // both CTAs' barriers initialised, TMEM allocated in both SMs
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  // shared::cluster addresses of the LEADER's barriers this CTA signals (pair: operand bytes of both CTAs are credited
  // to the leader's full barriers; the leader collects both CTAs' accumulator drains)
  [[maybe_unused]] const uint32_t full_ldr = PAIR ? mapa_u32(smem_u32(full_bar), 0) : 0u;
  [[maybe_unused]] const uint32_t tempty_ldr = PAIR ? mapa_u32(smem_u32(tempty_bar), 0) : 0u;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    int stage = 0, phase = 0;
    for (int t = tile0; t < total; t += tstep) {
      const int m0 = (t / n_tiles) * TM + m_off, nb = t % n_tiles;
      for (int kb = 0; kb < kb_count; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          if constexpr (PAIR) {
            if (leader) mbar_expect_tx(full_bar + stage, 2 * C::kStageBytes);
            const uint32_t bar = full_ldr + stage * 8;
            tma_load_2d_2sm(sa, &tmA, bar, kb * BK, m0);
            // the pair's B operand is CTA 0's rows followed by CTA 1's: GLU -> value half | gate half
            if constexpr (GLU) tma_load_2d_2sm(sb, &tmW, bar, kb * BK, (int)crank * p.N + nb * OUT_BN);
            else tma_load_2d_2sm(sb, &tmW, bar, kb * BK, nb * BN + (int)crank * (BN / 2));
          } else {
          mbar_expect_tx(full_bar + stage, C::kStageBytes);
          tma_load_2d(sa, &tmA, full_bar + stage, kb * BK, m0);
          if constexpr (GLU) {   // value half and gate half of [Wa;Wb] side by side in one B tile
            tma_load_2d(sb, &tmW, full_bar + stage, kb * BK, nb * OUT_BN);
            tma_load_2d(sb + C::kBBytes / 2, &tmW, full_bar + stage, kb * BK, p.N + nb * OUT_BN);
          } else {
            tma_load_2d(sb, &tmW, full_bar + stage, kb * BK, nb * BN);
          }
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues; pair: leader CTA only) ==========
    constexpr uint32_t idesc = umma_idesc_bf16(TM, BN);
    int stage = 0, phase = 0, it = 0;
    bool have = false;
    const bool tr = p.trace != nullptr && blockIdx.x == 0;
    long long w_acc = 0, w_full = 0, t_begin = tr ? clock64() : 0;
    for (int t = tile0; t < total; t += tstep, ++it) {
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      long long c0 = tr ? clock64() : 0;
      if constexpr (PAIR) mbar_wait_cluster(tempty_bar + acc, acc_phase ^ 1);   // both CTAs' epilogues have drained it
      else mbar_wait(tempty_bar + acc, acc_phase ^ 1);     // epilogue has drained this accumulator
      if (tr) w_acc += clock64() - c0;
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < kb_count; ++kb) {
        if (!have) { long long c1 = tr ? clock64() : 0; mbar_wait(full_bar + stage, phase); if (tr) w_full += clock64() - c1; }
        tc_fence_after();
        {   // probe the next slot; consumed after this stage's MMAs have been issued
          const int ns = (stage + 1 == C::kStages) ? 0 : stage + 1;
          have = mbar_test(full_bar + ns, (stage + 1 == C::kStages) ? (phase ^ 1) : phase);
        }
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + C::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // advance 32 bytes (16 bf16) inside the swizzle atom: +2 in the 16-byte-unit start address
            if constexpr (PAIR) umma_bf16_2sm(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (PAIR) {                         // multicast: the slot / accumulator barriers of BOTH CTAs
            umma_commit_2sm(empty_bar + stage, 0x3);
            if (kb == kb_count - 1) umma_commit_2sm(tfull_bar + acc, 0x3);
          } else {
          umma_commit(empty_bar + stage);             // frees the smem slot when these MMAs retire
          if (kb == kb_count - 1) umma_commit(tfull_bar + acc);   // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
    if (tr && lane == 0) { p.trace[0] = clock64() - t_begin; p.trace[1] = w_acc; p.trace[2] = w_full; p.trace[3] = it; }
  } else if (warp >= 4) {
    // ===================== epilogue (thread = output row; two warpgroups split the columns) ==============
    const int q = warp & 3;                           // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;                      // row inside the tile
    const int grp = (warp - 4) >> 2;                  // epilogue warpgroup
    const int et = threadIdx.x - 128 - grp * 128;     // 0..127 inside the warpgroup
    const bool elected = (et == 0);
    const int bar_id = 1 + grp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint32_t ring_phase = 0;                          // bit b = parity of the next completion of res_bar[b]
    int sub_cnt = 0;                                  // F1: running staging-buffer counter
    int it = 0;
    auto release_acc = [&](int a) {
      if constexpr (PAIR) mbar_arrive_cluster(tempty_ldr + a * 8);
      else mbar_arrive(tempty_bar + a);
    };
    for (int t = tile0; t < total; t += tstep, ++it) {
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      const int m0 = (t / n_tiles) * TM + m_off, n0 = (t % n_tiles) * OUT_BN;
      const uint32_t taddr = tmem_base + lane_base + acc * BN;

      // ---- per-tile parameters -> smem (previous tile's readers are past their last bar.sync); each warpgroup
      //      stages (and later reads) only the columns it owns
      if constexpr (RESID) {
        resid_stage_params<BN, 256>(sparam, threadIdx.x - 128, p.bias, n0, p.ln_mode, p.g1, p.b1, p.g2, p.b2);
      } else {
        constexpr int GCOLS = OUT_BN / NG;
        for (int ii = et; ii < GCOLS; ii += 128) {
          const int i = grp * GCOLS + ii;
          if constexpr (EPI == EPI_ARGMAX) sparam[i] = (p.bias && n0 + i < p.n_valid) ? p.bias[n0 + i] : 0.f;
          else
          sparam[i] = p.bias ? p.bias[n0 + i] : 0.f;
          if constexpr (GLU) sparam[OUT_BN + i] = p.bias ? p.bias[p.N + n0 + i] : 0.f;
        }
      }

      if constexpr (!RESID) {
        // ---------------- bf16 activations: 64-column sub-tiles through a 2-deep staging ring + TMA store
        const bool tre = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 128;
        long long e0 = tre ? clock64() : 0;
        mbar_wait(tfull_bar + acc, acc_phase);
        if (tre) { p.trace[4] += clock64() - e0; e0 = clock64(); }
        tc_fence_after();
        if constexpr (EPI == EPI_ARGMAX) {
          // row-wise (max, argmax) of this warpgroup's columns; tiles of the same rows on other CTAs are combined
          // through a 64-bit atomicMax on (order-preserving logit bits, ~column): ties go to the lowest column
          named_bar_sync(bar_id, 128);                 // sparam visible
          float best = -CUDART_INF_F;
          int best_col = 0;
#pragma unroll 1
          for (int ss = 0; ss < OUT_BN / 128; ++ss) {
            const int sub = grp * (OUT_BN / 128) + ss;
            uint32_t v[64];
            {
              uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
              uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
              tmem_ld32(taddr + sub * 64, v0);
              tmem_ld32(taddr + sub * 64 + 32, v1);
            }
            tmem_ld_wait();
            const float* bs = sparam + sub * 64;
            const int c0 = n0 + sub * 64;
#pragma unroll
            for (int e = 0; e < 64; ++e) {
              const float x = __uint_as_float(v[e]) + bs[e];
              if (c0 + e < p.n_valid && x > best) { best = x; best_col = c0 + e; }
            }
          }
          tc_fence_before();
          release_acc(acc);
          const int row = m0 + r;
          if (row < p.M && best > -CUDART_INF_F) {
            uint32_t u = __float_as_uint(best);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            atomicMax(p.keys + row, (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - (uint32_t)best_col));
          }
          named_bar_sync(bar_id, 128);                 // sparam may be restaged for the next tile
        } else {
#pragma unroll 1
        for (int ss = 0; ss < OUT_BN / (64 * NG); ++ss, ++sub_cnt) {
          const int sub = grp * (OUT_BN / (64 * NG)) + ss;   // 64-column sub-tile owned by this warpgroup
          // staging buffers: two per warpgroup (NG = 2) or one (NG = 4: its previous store must have left it)
          uint8_t* buf = ring + (NG == 2 ? grp * 2 + (sub_cnt & 1) : grp) * kBufBytes;
          if (elected) {
            long long e1 = tre ? clock64() : 0;
            if constexpr (NG == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
            if (tre) p.trace[5] += clock64() - e1;
          }
          named_bar_sync(bar_id, 128);                 // (also publishes sparam on the first sub-tile)
          uint32_t v[64];
          {
            uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
            uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
            tmem_ld32(taddr + sub * 64, v0);
            tmem_ld32(taddr + sub * 64 + 32, v1);
          }
          [[maybe_unused]] uint32_t g[GLU ? 64 : 1];
          if constexpr (GLU) {
            uint32_t (&g0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&g[0]);
            uint32_t (&g1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&g[32]);
            tmem_ld32(taddr + OUT_BN + sub * 64, g0);
            tmem_ld32(taddr + OUT_BN + sub * 64 + 32, g1);
          }
          tmem_ld_wait();
          const float* bs = sparam + sub * 64;
#pragma unroll
          for (int j = 0; j < 8; ++j) {                // 8 x 16-byte chunks = 64 bf16
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float x = __uint_as_float(v[8 * j + e]) + bs[8 * j + e];
              if constexpr (EPI == CFM_EPI_BIAS_SILU) x = silu_fast(x);
              if constexpr (EPI == CFM_EPI_BIAS_RELU) x = fmaxf(x, 0.f);
              if constexpr (GLU) x *= sigmoid_fast(__uint_as_float(g[8 * j + e]) + bs[OUT_BN + 8 * j + e]);
              f[e] = x;
            }
            *reinterpret_cast<uint4*>(buf + sw_off(r, j)) =
                make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (elected) {
            tma_store_2d(&tmC, buf, n0 + sub * 64, m0);   // rows >= M are clipped by the tensor map
            bulk_commit();
          }
        }
        tc_fence_before();
        release_acc(acc);
        if (tre) p.trace[6] += clock64() - e0;
        }
      } else {
        // ---------------- fp32 residual stream (+ fused LayerNorms): resid_epilogue.cuh
        constexpr int RG = C::kBufs / 2;                  // staging tiles per warpgroup
        uint8_t* gring = ring + grp * RG * kBufBytes;
        if (elected && !p.no_resid) resid_prefetch<BN, RG, 128, 2>(gring, res_bar + grp * RG, &tmR, n0, m0, grp);   // lands during the main loop
        mbar_wait(tfull_bar + acc, acc_phase);
        tc_fence_after();
        ResidParams rp{p.row_valid, p.y_row_valid, p.alpha, p.eps, p.ln_mode, p.M};
        ResidOpts ro;
        ro.no_residual = p.no_resid != 0;
        resid_ln_epilogue<BN, RG, 128, 2>(taddr, r, m0, n0, elected, bar_id, gring, res_bar + grp * RG, ring_phase, sparam, &tmC,
                                          &tmR, &tmY, rp, grp, 3, reinterpret_cast<float2*>(sparam + 5 * BN), -1, -1, ro);
        release_acc(acc);
      }
    }
    if (elected) bulk_wait_read<0>();   // the stores only have to be done READING shared memory before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();           // nobody frees TMEM / exits while the peer's MMAs may still touch it
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc_2sm<C::kTmemCols>(tmem_base);
    else tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN, int EPI, bool PAIR>
int launch_tc_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC, const CUtensorMap& tmR,
                   const CUtensorMap& tmY, const GemmParams& p, cudaStream_t st) {
  using C = Cfg<BN, EPI == CFM_EPI_RESIDUAL, PAIR, epi_groups<BN, EPI>()>;
  CFM_SMEM_OPT_IN((gemm_tc_kernel<BN, EPI, PAIR>), C::kSmemBytes);
  constexpr int OUT_BN = (EPI == CFM_EPI_BIAS_GLU) ? BN / 2 : BN;
  constexpr int TM = PAIR ? 2 * BM : BM;
  const int total = ((p.M + TM - 1) / TM) * (p.N / OUT_BN);
  const int slots = PAIR ? num_sms() / 2 : num_sms();
  const int grid = (total < slots ? total : slots) * (PAIR ? 2 : 1);
  CFM_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, EPI, PAIR>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st, PAIR ? 2 : 1, tmA,
                         tmW, tmC, tmR, tmY, p));
  CFM_LAUNCHED_K("gemm_tc");
  if (PAIR) count_variant("gemm_tc_pair");
  return 0;
}

// CTA pairs (256-row tiles) pay where the main loop is long: measured on the BASELINE shapes (tools/gemm_shapes_bench.py,
// L2 flushed) K=2048/N=512 54.3 -> 50.2 us, K=2048/N=256 37.9 -> 35.8, K=512/N=1536 35.8 -> 33.8; short-K tiles with a
// heavy epilogue lose (K=256/N=2048 + SiLU 29.6 -> 35.7, K=512/N=512 + residual 29.7 -> 31.7: the cluster-scope
// accumulator hand-back and the 2-CTA launch granularity cost more than the halved operand traffic saves).
// Warm cycle accounting of one CTA (tools/gemm_trace.py): K=512/N=2048 + SiLU 6.8 k -> 5.1 k cycles per 128 x 256 tile (the
// single-CTA MMA warp waits 28 % of its time for operands: 384 KB per tile against the ~70 B/clk a single SM's TMA
// delivers; the SiLU epilogue itself is 4.1 k cycles per tile, as long as the 4.1 k of MMAs), K=512/N=1536 6.6 k -> 4.8 k,
// K=2048/N=512 23.3 k -> 18.0 k (ideal 16.4 k).
// CFM_B200_GEMM_PAIR=0/1 forces the choice (measurement switch).
inline bool use_pair(int M, int N, int K, int epilogue) {
  static const bool off = env_is("CFM_B200_GEMM_PAIR", "0"), on = env_is("CFM_B200_GEMM_PAIR", "1");
  if (off || M <= BM) return false;
  if (on) return true;
  return K >= 1024 || (K >= 512 && N >= 1024 && epilogue != CFM_EPI_BIAS_GLU);
}

template <int BN, int EPI>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC, const CUtensorMap& tmR,
              const CUtensorMap& tmY, const GemmParams& p, cudaStream_t st) {
  if (p.pair) return launch_tc_impl<BN, EPI, true>(tmA, tmW, tmC, tmR, tmY, p, st);
  return launch_tc_impl<BN, EPI, false>(tmA, tmW, tmC, tmR, tmY, p, st);
}

// output-tile width: 256 when it divides N (fewer, fatter tiles), else 128
inline int pick_bn(int N, int epilogue) {
  if (epilogue == CFM_EPI_BIAS_GLU) return 256;          // 128 value + 128 gate columns
  return (N % 256 == 0) ? 256 : 128;
}

int make_2d(CUtensorMap* tm, bool f32, const void* base, int rows, int cols, int ld, int box_rows) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)ld * (f32 ? 4 : 2)};
  const uint32_t box[2] = {(uint32_t)(f32 ? 32 : 64), (uint32_t)box_rows};
  return f32 ? tc::make_tmap_f32(tm, base, 2, dims, str, box) : tc::make_tmap_bf16(tm, base, 2, dims, str, box);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool gemm_tc_supported(int lda, int ldc, int M, int N, int K, int dtype, int epilogue) {
  if (dtype != CFM_BF16 || tc::encode_tiled_fn() == nullptr) return false;
  if (K % BK != 0 || lda % 8 != 0 || M < 1) return false;
  if (N % 128 != 0) return false;
  if (epilogue == CFM_EPI_RESIDUAL ? (ldc % 4 != 0) : (ldc % 8 != 0)) return false;
  if (M < 64) return false;     // tiny streaming chunks: a 128-row MMA tile is >50 % padding, SIMT engine is used
  return true;
}

bool gemm_tc_ln_supported(int lda, int ldx, int ldy, int M, int N, int K, int dtype) {
  return gemm_tc_supported(lda, ldx, M, N, K, dtype, CFM_EPI_RESIDUAL) && N == 256 && ldy % 8 == 0;
}

int gemm_tc_init() {
  tc::encode_tiled_fn();
  return 0;
}

int gemm_tc_ln(const void* A, int lda, const void* W, const float* bias, float* X, int ldx, const float* residual, int M,
               int N, int K, float alpha, const uint8_t* row_valid, int ln_mode, const float* g1, const float* b1,
               const float* g2, const float* b2, void* Y, int ldy, const uint8_t* y_row_valid, float eps, int epilogue,
               void* Cact, int ldc, cudaStream_t st) {
  // common launcher: epilogue == RESIDUAL uses (X, residual, ln_*), otherwise (Cact, ldc)
  CFM_CHECK_ARG(aligned16(A) && aligned16(W), "cfm_gemm(tc): A/W must be 16-byte aligned");
  const int bn = pick_bn(N, epilogue);
  const int w_rows = (epilogue == CFM_EPI_BIAS_GLU) ? 2 * N : N;
  const bool pair = use_pair(M, N, K, epilogue);
  CUtensorMap tmA, tmW, tmC, tmR, tmY;
  int rc;
  if ((rc = make_2d(&tmA, false, A, M, K, lda, BM)) != 0) return rc;
  if ((rc = make_2d(&tmW, false, W, w_rows, K, K, (epilogue == CFM_EPI_BIAS_GLU) ? 128 : (pair ? bn / 2 : bn))) != 0) return rc;
  GemmParams p{};
  p.pair = pair ? 1 : 0;
  if (const char* e = getenv("CFM_B200_GEMM_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  p.bias = bias; p.row_valid = row_valid; p.y_row_valid = y_row_valid;
  p.g1 = g1; p.b1 = b1; p.g2 = g2; p.b2 = b2;
  p.alpha = alpha; p.eps = eps; p.M = M; p.N = N; p.K = K; p.ln_mode = ln_mode;
  if (epilogue == CFM_EPI_RESIDUAL) {
    p.no_resid = residual == nullptr ? 1 : 0;
    if (residual == nullptr) residual = X;           // (tensor map only; never loaded)
    CFM_CHECK_ARG(aligned16(X) && aligned16(residual), "cfm_gemm(tc): X/residual must be 16-byte aligned");
    if ((rc = make_2d(&tmC, true, X, M, N, ldx, BM)) != 0) return rc;
    if ((rc = make_2d(&tmR, true, residual, M, N, ldx, BM)) != 0) return rc;
    tmY = tmC;
    if (ln_mode != 0) {
      CFM_CHECK_ARG(N == bn, "cfm_gemm_ln(tc): fused LayerNorm needs the whole row in one tile (N=%d, tile %d)", N, bn);
      CFM_CHECK_ARG(aligned16(Y) && ldy % 8 == 0, "cfm_gemm_ln(tc): Y must be 16-byte aligned with ldy %% 8 == 0");
      if ((rc = make_2d(&tmY, false, Y, M, N, ldy, BM)) != 0) return rc;
    }
    if (bn == 256) return launch_tc<256, CFM_EPI_RESIDUAL>(tmA, tmW, tmC, tmR, tmY, p, st);
    return launch_tc<128, CFM_EPI_RESIDUAL>(tmA, tmW, tmC, tmR, tmY, p, st);
  }
  CFM_CHECK_ARG(aligned16(Cact), "cfm_gemm(tc): C must be 16-byte aligned");
  if ((rc = make_2d(&tmC, false, Cact, M, N, ldc, BM)) != 0) return rc;
  tmR = tmC; tmY = tmC;
  if (bn == 256) {
    switch (epilogue) {
      case CFM_EPI_BIAS: return launch_tc<256, CFM_EPI_BIAS>(tmA, tmW, tmC, tmR, tmY, p, st);
      case CFM_EPI_BIAS_SILU: return launch_tc<256, CFM_EPI_BIAS_SILU>(tmA, tmW, tmC, tmR, tmY, p, st);
      case CFM_EPI_BIAS_RELU: return launch_tc<256, CFM_EPI_BIAS_RELU>(tmA, tmW, tmC, tmR, tmY, p, st);
      default: return launch_tc<256, CFM_EPI_BIAS_GLU>(tmA, tmW, tmC, tmR, tmY, p, st);
    }
  }
  switch (epilogue) {
    case CFM_EPI_BIAS: return launch_tc<128, CFM_EPI_BIAS>(tmA, tmW, tmC, tmR, tmY, p, st);
    case CFM_EPI_BIAS_RELU: return launch_tc<128, CFM_EPI_BIAS_RELU>(tmA, tmW, tmC, tmR, tmY, p, st);
    default: return launch_tc<128, CFM_EPI_BIAS_SILU>(tmA, tmW, tmC, tmR, tmY, p, st);
  }
}

// P = A W^T + bias is never written: per-row (max, argmax) over the V = N valid columns into `keys` (zeroed by the caller)
int gemm_tc_argmax(const void* A, int lda, const void* W, const float* bias, int M, int V, int K, unsigned long long* keys,
                   cudaStream_t st) {
  CFM_CHECK_ARG(aligned16(A) && aligned16(W), "cfm_ctc_argmax(tc): x/W must be 16-byte aligned");
  const int n_pad = (V + 255) / 256 * 256;
  CUtensorMap tmA, tmW;
  int rc;
  if ((rc = make_2d(&tmA, false, A, M, K, lda, BM)) != 0) return rc;
  const bool pair = use_pair(M, n_pad, K, EPI_ARGMAX);
  if ((rc = make_2d(&tmW, false, W, V, K, K, pair ? 128 : 256)) != 0) return rc;     // rows >= V of the last block: zero-filled
  GemmParams p{};
  p.pair = pair ? 1 : 0;
  p.bias = bias; p.M = M; p.N = n_pad; p.K = K; p.keys = keys; p.n_valid = V;
  return launch_tc<256, EPI_ARGMAX>(tmA, tmW, tmA, tmA, tmA, p, st);
}

int gemm_tc(const void* A, int lda, const void* W, const float* bias, void* C, int ldc, int M, int N, int K, int dtype,
            int epilogue, const float* residual, float alpha, const uint8_t* row_valid, cudaStream_t st) {
  if (epilogue == CFM_EPI_RESIDUAL)
    return gemm_tc_ln(A, lda, W, bias, (float*)C, ldc, residual, M, N, K, alpha, row_valid, 0, nullptr, nullptr, nullptr,
                      nullptr, nullptr, 0, nullptr, 0.f, epilogue, nullptr, 0, st);
  return gemm_tc_ln(A, lda, W, bias, nullptr, 0, nullptr, M, N, K, alpha, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
                    nullptr, 0, nullptr, 0.f, epilogue, C, ldc, st);
}

}  // namespace cfm
