// Fused macaron feed-forward for sm_100a (d = 256):
//     X += alpha * ( W2 . silu(W1 . y + b1) + b2 )        [+ the LayerNorm(s) that follow, in the same kernel]
// replacing w_1 -> SiLU -> w_2 of feedforward.py:16-21 plus encoder_layer.py:58-59 / 69-70,56 of the reference.
// The (tokens x 2048) hidden activation never leaves the SM: per 128-token tile the hidden dimension is walked in
// chunks of 128 units, two chunks (a pair) per G1 job,
//     S_p = y . W1_p^T        tcgen05.mma M=128 N=256 K=256   -> TMEM S (256 columns = chunks 2p, 2p+1)
//     H_c = silu(S_c + b1_c)  8 epilogue warps: tcgen05.ld -> tanh.approx -> bf16 -> swizzled smem (A operand),
//                             handed over in two 64-column halves (= the k atoms of G2)
//     Y  += H_c . W2_c^T      tcgen05.mma M=128 N=256 K=128   -> TMEM Y (256 columns, lives for the whole tile)
// and the tensor pipe runs  G1(p+1) G2(2p) G2(2p+1)  so that the SiLU of pair p+1 overlaps the G2s of pair p and
// G1(p+2).  Every MMA has N = 256: an N = 128 MMA with both operands in shared memory reads 128 B/clk, the whole SM
// budget, and was measured 50 % slower than nominal next to the TMA / SiLU traffic.  Weights stream through a ring
// of 32 KB TMA pieces (256 rows x 64 k, 128-byte swizzle); the input tile y (64 KB) is loaded once.
// The final epilogue is the shared residual/LayerNorm epilogue (resid_epilogue.cuh) on the Y accumulator; its
// staging rings alias the H buffers and the input tile, which are dead by then.
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-11: epilogue (two warpgroups,
//   each owning 32 columns of both halves of a chunk, and 128 of the 256 output columns in the final epilogue).
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include "resid_epilogue.cuh"
#include <stdlib.h>

namespace cfm {
namespace {

using namespace tc;

constexpr int D = 256;          // model dim = K of GEMM1 = N of GEMM2
constexpr int HC = 128;         // hidden units per chunk
constexpr int BM = 128;
constexpr int kAtom = 16384;    // one swizzle atom tile: 128 rows x 64 k bf16
constexpr int kPiece = 32768;   // one weight piece = one ring slot = 512 tensor-pipe cycles of MMA work per barrier wait
                                //   G1: W1 rows [256 p,+256) x 64 k   (4 MMAs N=256)
                                //   G2: W2 rows [0,256) x 64 hidden-k (4 MMAs N=256)
constexpr int NST = 3;          // weight ring depth (single-CTA kernel: 3 x 32 KB)
constexpr int kSlotPair = kPiece / 2;   // CTA-pair kernel: each CTA holds HALF of every weight piece (its 128 of the 256
constexpr int NST_PAIR = 6;             // rows) -> the same 96 KB hold a 6-deep ring
static_assert(NST * kPiece == NST_PAIR * kSlotPair, "both kernels share one shared-memory layout");
constexpr int kMaxStages = NST_PAIR;
constexpr int kABytes = BM * D * 2;            // 64 KB
constexpr int kHBytes = BM * HC * 2;           // 32 KB per H buffer
#ifndef CFM_FFN_SILU_GROUPS
#define CFM_FFN_SILU_GROUPS 2
#endif
constexpr int kSiluGroups = CFM_FFN_SILU_GROUPS;   // SiLU-stage warpgroups (2 or 4)
constexpr int kSiluCols = HC / kSiluGroups;        // hidden columns per warpgroup and chunk: half of them in each 64-column
constexpr int kHalfCols = kSiluCols / 2;           // half (= k atom of G2) of the chunk
static_assert(kSiluGroups == 2 || kSiluGroups == 4, "SiLU stage: 2 or 4 warpgroups");
constexpr int kSiluThreads = 128 * kSiluGroups;
constexpr int kThreads = 128 + kSiluThreads;
constexpr int kParamFloats = 2 * HC;           // double-buffered b1 chunk (the residual-epilogue parameters alias the
                                               // input tile, which is dead by the time they are needed)
constexpr int kSmemBytes = kABytes + 2 * kHBytes + NST * kPiece + kParamFloats * 4 + 512;
static_assert(5 * D * 4 <= kABytes, "residual parameters alias the input tile");
static_assert(kSmemBytes <= 232448, "smem budget");
static_assert(2 * kHBytes == 4 * kBufBytes, "residual staging ring aliases the two H buffers");

struct FfnStage {                  // one feed-forward module
  const float* b1;
  const float* b2;
  const float* g1; const float* be1; const float* g2; const float* be2;
  float alpha;
  int ln_mode;
};
struct FfnParams {
  FfnStage st[2];
  int n_stages;                    // 2: two modules chained on the same tile (X and y of the first stay on chip)
  const uint8_t* y_row_valid;      // row mask of the LAST stage's y
  const float* bp;                 // projection tail: P = y_last . Wp^T + bp (the QKV projection), NP256 blocks of 256 columns
  int np_blocks;                   // 0: no projection
  float eps;
  int M, F;
  long long* trace;   // optional per-event clock64 timestamps of CTA 0 (tools/ffn_trace.py); nullptr in production
};

// job jx of a tile.  G1(p) computes S for the PAIR p of hidden chunks (N = 256: with N = 128 a tcgen05.mma reads
// 128 B of shared memory per cycle, the whole SM budget, and every TMA / SiLU store slows it down), G2(c) consumes
// chunk c.  Order:  G1(0) | G1(1) G2(0) G2(1) | G1(2) G2(2) G2(3) | ... | G1(NP-1) G2(2NP-4) G2(2NP-3) | G2(2NP-2) G2(2NP-1)
// so that the SiLU of pair p overlaps G1(p+1) and the G2s of pair p-1.  idx = pair (G1) or chunk (G2).
__device__ __forceinline__ void job_of(int jx, int NP, bool& g1, int& idx) {
  if (jx == 0) { g1 = true; idx = 0; return; }
  const int q = (jx - 1) / 3, r = (jx - 1) % 3;
  if (q < NP - 1) {
    if (r == 0) { g1 = true; idx = q + 1; } else { g1 = false; idx = 2 * q + r - 1; }
  } else {
    g1 = false; idx = 2 * (NP - 1) + (jx - 1 - 3 * (NP - 1));
  }
}

// CL = thread-block cluster size along M (1 or 2).  With CL == 2 the two CTAs of a cluster work on adjacent
// 128-token tiles in lock-step and share every weight piece: each CTA fetches half of the piece's rows and
// TMA-multicasts it into both CTAs' rings, halving the L2 -> SM weight traffic (2 MB per tile otherwise).
//
// PAIR: tcgen05 cta_group::2.  The two CTAs of a cluster own adjacent 128-token tiles; every MMA spans both SMs (M = 256),
// issued by the leader CTA, and takes rows [0,128) of its B operand from the leader's shared memory and rows [128,256)
// from the peer's.  Per SM this halves the weight fill (TMA) and the B-operand reads -- the shared-memory port is what
// holds the single-CTA kernel at ~50 % tensor-pipe activity -- and the ring gets twice as deep.  Barriers the MMA issuer
// waits on live in the leader and collect warp-aggregated remote arrivals from both CTAs; barriers the other warps wait
// on are signalled in both CTAs by multicast commits.
template <int CL, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmA,    // y  (M, 256) bf16, box 64 x 128
                 const __grid_constant__ CUtensorMap tmW1,   // W1 (F, 256) bf16, box 64 x 128
                 const __grid_constant__ CUtensorMap tmW2,   // W2 (256, F) bf16, box 64 x 128
                 const __grid_constant__ CUtensorMap tmW1b,  // second module of a chain (same shapes)
                 const __grid_constant__ CUtensorMap tmW2b,
                 const __grid_constant__ CUtensorMap tmX,    // X  (M, 256) fp32 store, box 32 x 128
                 const __grid_constant__ CUtensorMap tmR,    // residual load (same tensor as X)
                 const __grid_constant__ CUtensorMap tmY,    // y out (M, 256) bf16 store, box 64 x 128
                 const __grid_constant__ CUtensorMap tmWp,   // projection weight (Np, 256) bf16, box 64 x 128
                 const __grid_constant__ CUtensorMap tmP,    // projection output (M, Np) bf16 store, box 64 x 128
                 const FfnParams p) {
  // 1024-byte alignment is what SWIZZLE_128B tiles need; keeping `smem` a plain shared-space array (no integer
  // round-up) lets ptxas emit LDS/STS instead of generic LD.E/ST.E for every epilogue access
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sH = sA + kABytes;                 // 2 x 32 KB; also the residual staging ring (4 x 16 KB)
  uint8_t* sW = sH + 2 * kHBytes;             // NST x 32 KB
  float* sparam = reinterpret_cast<float*>(sA);   // residual-epilogue parameters alias the input tile, which is dead after
                                                  // the last MMA of the tile and is only ever written by this CTA's own TMA
  float* sb1 = reinterpret_cast<float*>(sW + NST * kPiece);   // [2][HC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb1 + 2 * HC);
  constexpr int kNst = PAIR ? NST_PAIR : NST;
  constexpr int kSlot = PAIR ? kSlotPair : kPiece;
  static_assert(!PAIR || CL == 1, "the pair kernel does not multicast");
  uint64_t* w_full = bars;                    // [kNst]
  uint64_t* w_empty = w_full + kMaxStages;    // [kNst]
  uint64_t* a_full = w_empty + kMaxStages;    // [1]
  uint64_t* s_full = a_full + 1;              // [1]  S accumulator of a chunk pair complete (MMA commit)
  uint64_t* s_empty = s_full + 1;             // [1]  S read by all SiLU threads
  uint64_t* h_full = s_empty + 1;             // [2][2]  64-column half of H[b] written by the SiLU threads
  uint64_t* h_empty = h_full + 4;             // [2]  G2 finished reading H[b] (MMA commit)
  uint64_t* y_full = h_empty + 2;             // [1]  all MMAs of the tile complete
  uint64_t* tile_done = y_full + 1;           // [1]  final epilogue of the tile finished (256 arrivals)
  uint64_t* res_bar = tile_done + 1;          // [2 groups][4]
  uint64_t* a_ready = res_bar + 8;            // [1]  chain: y of the first module written into sA (256 arrivals)
  uint64_t* y_ready = a_ready + 1;            // [1]  projection tail: y of the last module written into sH (256 arrivals)
  uint64_t* q_full = y_ready + 1;             // [2]  projection accumulator complete (MMA commit)
  uint64_t* q_empty = q_full + 2;             // [2]  projection accumulator read out (256 arrivals)
  uint64_t* pair_done = q_empty + 2;          // [1]  PAIR: both CTAs' final epilogues finished (leader's instance)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pair_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // per-CTA wall-clock stamps (globaltimer, ns) when tracing: [16*64 + 4*cta + {0: entry, 1: prologue done, 2: all warps done}]
  auto gstamp = [&](int slot) {
    if (p.trace && threadIdx.x == 0) {
      unsigned long long tns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
      p.trace[16 * 64 + 4 * blockIdx.x + slot] = (long long)tns;
    }
  };
  gstamp(0);
  // every CTA of a cluster runs the same number of tiles (phantom tiles past M are fully out of bounds: TMA
  // zero-fills their loads and clips their stores) so that the shared weight ring stays in lock-step
  // (PAIR: m_tiles counts 256-row pair tiles; the peer of the last pair may lie completely past M)
  const int m_tiles = PAIR ? ((p.M + BM - 1) / BM + 1) / 2 : ((p.M + BM - 1) / BM + CL - 1) / CL * CL;
  const int NC = p.F / HC;
  const int NP = NC / 2;                     // pairs of hidden chunks (F % 256 == 0)
  const int n_jobs = 3 * NP;
  const uint32_t crank = (CL > 1 || PAIR) ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (1u << CL) - 1u;
  const bool leader = !PAIR || crank == 0;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TM = PAIR ? 2 * BM : BM;
  const int m_off = PAIR ? (int)crank * BM : 0;
  // arrivals on the barriers the MMA issuer waits on: per thread (single CTA) or one remote arrive per warp on the
  // leader's instance (pair)
  constexpr uint32_t kArrSilu = PAIR ? 2 * (kSiluThreads / 32) : kSiluThreads;
  constexpr uint32_t kArr256 = PAIR ? 2 * 8 : 256;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    prefetch_tmap(&tmX); prefetch_tmap(&tmR); prefetch_tmap(&tmY);
    if (p.n_stages > 1) { prefetch_tmap(&tmW1b); prefetch_tmap(&tmW2b); }
    if (p.np_blocks > 0) { prefetch_tmap(&tmWp); prefetch_tmap(&tmP); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kNst; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, CL); }
    mbar_init(a_full, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, kArrSilu);
    for (int s = 0; s < 2; ++s) {
      mbar_init(h_full + 2 * s, kArrSilu); mbar_init(h_full + 2 * s + 1, kArrSilu); mbar_init(h_empty + s, 1);
    }
    mbar_init(y_full, 1);
    mbar_init(tile_done, 256);
    for (int s = 0; s < 8; ++s) mbar_init(res_bar + s, 1);
    mbar_init(a_ready, kArr256); mbar_init(y_ready, kArr256);
    for (int s = 0; s < 2; ++s) { mbar_init(q_full + s, 1); mbar_init(q_empty + s, kArr256); }
    mbar_init(pair_done, kArr256);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_2sm<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1 || PAIR) cluster_sync_all();   // peer barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  pdl_wait();
  gstamp(1);
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_y = tmem_base + 256;
  // shared::cluster address of the LEADER's copy of a barrier = local shared::cta address + ldr_off
  [[maybe_unused]] const uint32_t ldr_off = PAIR ? mapa_u32(smem_u32(bars), 0) - smem_u32(bars) : 0u;
  // arrive on a barrier the MMA issuer waits on
  auto arrive_mma = [&](uint64_t* bar) {
    if constexpr (PAIR) {
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(bar) + ldr_off);
    } else {
      mbar_arrive(bar);
    }
  };
  // the MMA issuer's waits / probes on those barriers (cluster-scope acquire when the peer CTA arrives on them)
  auto wait_mma = [&](uint64_t* bar, uint32_t parity) {
    if constexpr (PAIR) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
  };
  auto test_mma = [&](uint64_t* bar, uint32_t parity) {
    if constexpr (PAIR) return mbar_test_cluster(bar, parity); else return mbar_test(bar, parity);
  };

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) =====================
    int stage = 0, phase = 0, it = 0;
    for (int t = tile0; t < m_tiles; t += tstep, ++it) {   // gridDim.x is a multiple of CL
      const int m0 = t * TM + m_off;
      if (it > 0) mbar_wait(tile_done, (it - 1) & 1);     // sA / sH / ring of the previous tile are dead
      if (elect_one()) {
        if constexpr (PAIR) {
          if (leader) mbar_expect_tx(a_full, 2 * kABytes);
#pragma unroll
          for (int ka = 0; ka < D / 64; ++ka) tma_load_2d_2sm(sA + ka * kAtom, &tmA, smem_u32(a_full) + ldr_off, ka * 64, m0);
        } else {
        mbar_expect_tx(a_full, kABytes);
#pragma unroll
        for (int ka = 0; ka < D / 64; ++ka) tma_load_2d(sA + ka * kAtom, &tmA, a_full, ka * 64, m0);
        }
      }
      __syncwarp();
      for (int sg = 0; sg < p.n_stages; ++sg) {
      for (int jx = 0; jx < n_jobs; ++jx) {
        bool g1; int c;
        job_of(jx, NP, g1, c);
        const int n_pc = g1 ? 4 : 2;
        for (int pc = 0; pc < n_pc; ++pc) {
          mbar_wait(w_empty + stage, phase ^ 1);      // CL == 2: both CTAs have released this slot
          if (elect_one()) {
            uint8_t* dst = sW + stage * kSlot;
            if (!PAIR || leader) mbar_expect_tx(w_full + stage, kPiece);
            // G1: W1 rows [256 c, +256) x k [64 pc, +64);  G2: W2 rows [0,256) x hidden k [128 c + 64 pc, +64)
            const CUtensorMap* tm = g1 ? (sg ? &tmW1b : &tmW1) : (sg ? &tmW2b : &tmW2);
            const int col = g1 ? pc * 64 : c * HC + pc * 64;
            const int row = g1 ? c * 256 : 0;
            if constexpr (PAIR) {      // this CTA's 128 of the piece's 256 rows; bytes credited to the leader's barrier
              tma_load_2d_2sm(dst, tm, smem_u32(w_full + stage) + ldr_off, col, row + (int)crank * 128);
            } else if constexpr (CL == 1) {
              tma_load_2d(dst, tm, w_full + stage, col, row);
              tma_load_2d(dst + kAtom, tm, w_full + stage, col, row + 128);
            } else {           // each CTA fetches 256 / CL of the 256 rows and multicasts them to the whole cluster
              tma_load_2d_mc(dst + crank * (kPiece / CL), tm, w_full + stage, col, row + crank * (256 / CL), kMask);
            }
          }
          __syncwarp();
          if (++stage == kNst) { stage = 0; phase ^= 1; }
        }
      }
      }
      // projection tail: Wp rows [256 blk, +256) x k [64 kc, +64)
      for (int pc = 0; pc < 4 * p.np_blocks; ++pc) {
        mbar_wait(w_empty + stage, phase ^ 1);
        if (elect_one()) {
          uint8_t* dst = sW + stage * kSlot;
          if (!PAIR || leader) mbar_expect_tx(w_full + stage, kPiece);
          const int col = (pc & 3) * 64, row = (pc >> 2) * 256;
          if constexpr (PAIR) {
            tma_load_2d_2sm(dst, &tmWp, smem_u32(w_full + stage) + ldr_off, col, row + (int)crank * 128);
          } else if constexpr (CL == 1) {
            tma_load_2d(dst, &tmWp, w_full + stage, col, row);
            tma_load_2d(dst + kAtom, &tmWp, w_full + stage, col, row + 128);
          } else {
            tma_load_2d_mc(dst + crank * (kPiece / CL), &tmWp, w_full + stage, col, row + crank * (256 / CL), kMask);
          }
        }
        __syncwarp();
        if (++stage == kNst) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues; pair: leader CTA only) ==========
    constexpr uint32_t idesc = umma_idesc_bf16(TM, 256);    // G1: 256 hidden units of a chunk pair, G2: 256 outputs
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
      if constexpr (PAIR) umma_bf16_2sm(d, da, db, idesc, acc); else umma_bf16(d, da, db, idesc, acc);
    };
    auto commit = [&](uint64_t* bar) {           // barrier other warps wait on: both CTAs' instances in pair mode
      if constexpr (PAIR) umma_commit_2sm(bar, 0x3); else umma_commit(bar);
    };
    auto commit_slot = [&](uint64_t* bar) {
      if constexpr (PAIR) umma_commit_2sm(bar, 0x3);
      else if constexpr (CL == 1) umma_commit(bar);
      else umma_commit_mc(bar, kMask);
    };
    int stage = 0, phase = 0, it = 0;
    bool have = false;   // w_full of the current slot already seen complete by the probe issued before the previous MMAs
    uint32_t n_se = 0, n_hf0 = 0, n_hf1 = 0;
    for (int t = tile0; t < m_tiles; t += tstep, ++it) {   // gridDim.x is a multiple of CL
      if (it > 0) {                                        // Y accumulator(s) drained by the previous epilogue(s)
        if constexpr (PAIR) mbar_wait_cluster(pair_done, (it - 1) & 1); else mbar_wait(tile_done, (it - 1) & 1);
      }
      const uint32_t a_addr = smem_u32(sA), h_addr = smem_u32(sH);
      bool job_ready = false;           // the next job's S/H barrier was already seen complete
      auto next_slot_probe = [&]() {
        const int ns = (stage + 1 == kNst) ? 0 : stage + 1;
        have = mbar_test(w_full + ns, (stage + 1 == kNst) ? (phase ^ 1) : phase);
      };
      auto advance = [&]() { if (++stage == kNst) { stage = 0; phase ^= 1; } };
      for (int sg = 0; sg < p.n_stages; ++sg) {
      // input tile: from TMA (first module) or written by the first module's epilogue, which also left X / alpha in Y
      if (p.trace && blockIdx.x == 0 && lane == 0 && sg == 0) p.trace[10 * 64 + 0] = clock64();
      if (sg == 0) mbar_wait(a_full, it & 1); else wait_mma(a_ready, it & 1);
      tc_fence_after();
      if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[10 * 64 + 1 + sg] = clock64();
      const uint32_t y_acc0 = sg;                         // chained module: accumulate onto the parked residual
      // The issuing warp runs in lock-step with the tensor pipe (it accepts only a few MMAs ahead), so every cycle
      // between two issue blocks is an idle pipe cycle.  The job sequence is therefore written out without any per-job
      // decoding, and the barrier of the NEXT issue block (ring slot or S/H hand-over) is probed non-blockingly
      // before the current block's MMAs are issued.
      // probe used while issuing the last piece of a job: is the barrier of the following job complete?
      auto probe_g1 = [&]() { return test_mma(s_empty, (n_se & 1) ^ 1); };
      auto probe_g2 = [&](int c) { const int b = c & 1; return test_mma(h_full + 2 * b, (b ? n_hf1 : n_hf0) & 1); };

      auto do_g1 = [&](int pr, int next_kind, int next_c) {    // next_kind: 1 = G1, 2 = G2, 0 = none
        if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[0 * 64 + pr] = clock64();
        if (!job_ready) wait_mma(s_empty, (n_se & 1) ^ 1);         // SiLU stage has read S of the previous pair
        if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[1 * 64 + pr] = clock64();
        ++n_se;
        tc_fence_after();
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {                             // pc = 64-wide k atom of the input tile
          if (!have) mbar_wait(w_full + stage, phase);
          tc_fence_after();
          next_slot_probe();
          if (pc == 3) job_ready = (next_kind == 1) ? probe_g1() : (next_kind == 2 ? probe_g2(next_c) : false);
          if (elect_one()) {
            const uint64_t da = umma_desc_sw128(a_addr + pc * kAtom);
            const uint64_t db = umma_desc_sw128(smem_u32(sW + stage * kSlot));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tmem_base, da + 2 * k, db + 2 * k, (pc | k) != 0);
            commit_slot(w_empty + stage);
            if (pc == 3) commit(s_full);
          }
          __syncwarp();
          advance();
        }
      };
      auto do_g2 = [&](int c, int next_kind, int next_c, bool last) {
        const int b = c & 1;
        uint32_t& n_hf = b ? n_hf1 : n_hf0;
        if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[6 * 64 + c] = clock64();
        if (!job_ready) wait_mma(h_full + 2 * b, n_hf & 1);         // first half of H[b] written (and fenced)
        if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[7 * 64 + c] = clock64();
        tc_fence_after();
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {                             // pc = 64-wide k atom of the hidden chunk
          if (pc == 1) {                                             // second half: lands while the first atom's MMAs run
            if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[8 * 64 + c] = clock64();
            wait_mma(h_full + 2 * b + 1, n_hf & 1);
            if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[9 * 64 + c] = clock64();
          }
          if (!have) mbar_wait(w_full + stage, phase);
          tc_fence_after();
          next_slot_probe();
          if (pc == 1) job_ready = (next_kind == 1) ? probe_g1() : (next_kind == 2 ? probe_g2(next_c) : false);
          if (elect_one()) {
            const uint64_t da = umma_desc_sw128(h_addr + b * kHBytes + pc * kAtom);
            const uint64_t db = umma_desc_sw128(smem_u32(sW + stage * kSlot));
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(tmem_y, da + 2 * k, db + 2 * k, (y_acc0 | c | pc | k) != 0);
            commit_slot(w_empty + stage);
            if (pc == 1) {
              commit(h_empty + b);
              if (last) commit(y_full);
            }
          }
          __syncwarp();
          advance();
        }
        ++n_hf;
      };
      job_ready = false;
      do_g1(0, NP > 1 ? 1 : 2, 0);
      for (int pr = 0; pr + 1 < NP; ++pr) {
        do_g1(pr + 1, 2, 2 * pr);
        do_g2(2 * pr, 2, 2 * pr + 1, false);
        do_g2(2 * pr + 1, pr + 2 < NP ? 1 : 2, 2 * pr + 2, false);
      }
      do_g2(NC - 2, 2, NC - 1, false);
      do_g2(NC - 1, 0, 0, true);
      if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[10 * 64 + 3 + sg] = clock64();
      }
      // projection tail: P_blk = y . Wp_blk^T, accumulators alternate between the S and the Y columns
      if (p.np_blocks > 0) {
        wait_mma(y_ready, it & 1);                       // y of the last module is in sH, the Y columns are dead
        tc_fence_after();
        if (p.trace && blockIdx.x == 0 && lane == 0) p.trace[10 * 64 + 5] = clock64();
        for (int blk = 0; blk < p.np_blocks; ++blk) {
          const int ab = blk & 1;
          // accumulator `ab` is used by blocks ab, ab+2, ...: wait until the epilogue has read out block blk-2
          if (blk >= 2) { wait_mma(q_empty + ab, (it * ((p.np_blocks + 1 - ab) >> 1) + (blk >> 1) - 1) & 1); tc_fence_after(); }
          for (int kc = 0; kc < 4; ++kc) {
            if (!have) mbar_wait(w_full + stage, phase);
            tc_fence_after();
            next_slot_probe();
            if (elect_one()) {
              const uint64_t da = umma_desc_sw128(h_addr + kc * kAtom);
              const uint64_t db = umma_desc_sw128(smem_u32(sW + stage * kSlot));
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(ab ? tmem_y : tmem_base, da + 2 * k, db + 2 * k, (kc | k) != 0);
              commit_slot(w_empty + stage);
              if (kc == 3) commit(q_full + ab);
            }
            __syncwarp();
            advance();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - grp * 128;
    const bool elected = (et == 0);
    const int bar_id = 1 + grp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint32_t ring_phase = 0;
    uint32_t n_sf = 0, n_he0 = 0, n_he1 = 0;
    int it = 0;
    for (int t = tile0; t < m_tiles; t += tstep, ++it) {   // gridDim.x is a multiple of CL
      const int m0 = t * TM + m_off;
      if (it > 0 && grp >= 2) mbar_wait(tile_done, (it - 1) & 1);   // H buffers double as group 0's staging ring
      for (int sg = 0; sg < p.n_stages; ++sg) {
      const FfnStage& fs = p.st[sg];
      // ---- SiLU stage: S (chunk pair) -> H[0], H[1].  The b1 slice of the next pair is fetched into registers while
      //      this pair is being processed, so its L2 latency never sits between two pairs.
      // et < kSiluCols: the hidden columns this warpgroup reads (kHalfCols of each 64-column half)
      const int bcol = (et / kHalfCols) * 64 + grp * kHalfCols + (et % kHalfCols);
      float b1_n0 = 0.f, b1_n1 = 0.f;
      if (et < kSiluCols) { b1_n0 = __ldg(fs.b1 + bcol); b1_n1 = __ldg(fs.b1 + HC + bcol); }
      for (int pr = 0; pr < NP; ++pr) {
        named_bar_sync(bar_id, 128);                      // the group has finished reading the previous pair's b1
        if (et < kSiluCols) {
          sb1[bcol] = b1_n0;
          sb1[HC + bcol] = b1_n1;
          if (pr + 1 < NP) { b1_n0 = __ldg(fs.b1 + (2 * pr + 2) * HC + bcol); b1_n1 = __ldg(fs.b1 + (2 * pr + 3) * HC + bcol); }
        }
        named_bar_sync(bar_id, 128);
        if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[2 * 64 + 2 * pr] = clock64();
        mbar_wait(s_full, n_sf & 1);                      // S holds both chunks of the pair
        ++n_sf;
        tc_fence_after();
        // chain: the last G1 of the first module has retired, so the input tile is dead: start fetching the residual
        // rows of its epilogue into it now (2 x 16 KB per warpgroup), while the remaining G2 jobs run
        if (pr == NP - 1 && sg + 1 < p.n_stages && elected && grp < 2)
          resid_prefetch<D, 2, 128, 2>(sA + grp * 2 * kBufBytes, res_bar + grp * 4, &tmR, 0, m0, grp);
        if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[3 * 64 + 2 * pr] = clock64();
        // S is read out completely first so that the next pair's G1 can start; v[b][half] = this warpgroup's kHalfCols
        // columns of 64-column half `half` (= k atom of G2) of chunk 2 pr + b
        uint32_t v[2][2][kHalfCols];
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int half = 0; half < 2; ++half) tmem_ld(tmem_base + lane_base + b * HC + half * 64 + grp * kHalfCols, v[b][half]);
        tmem_ld_wait();
        tc_fence_before();
        arrive_mma(s_empty);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          uint32_t& n_he = b ? n_he1 : n_he0;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const float* bs = sb1 + b * HC + half * 64 + grp * kHalfCols;
            uint4 pk[kHalfCols / 8];
#pragma unroll
            for (int j = 0; j < kHalfCols / 8; ++j) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = silu_fast(__uint_as_float(v[b][half][8 * j + e]) + bs[8 * j + e]);
              pk[j] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
            if (half == 0) {
              if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[5 * 64 + 2 * pr + b] = clock64();
              mbar_wait(h_empty + b, (n_he & 1) ^ 1);       // G2 of two chunks ago finished reading H[b]
              ++n_he;
            }
            uint8_t* hb = sH + b * kHBytes + half * kAtom;  // this group's 16-byte chunks of the half's swizzle atom
#pragma unroll
            for (int j = 0; j < kHalfCols / 8; ++j) *reinterpret_cast<uint4*>(hb + sw_off(r, grp * (kHalfCols / 8) + j)) = pk[j];
            fence_proxy_async_smem();
            arrive_mma(h_full + 2 * b + half);
          }
          if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[4 * 64 + 2 * pr + b] = clock64();
        }
      }
      // ---- epilogue on Y: warpgroups 0 and 1 take 128 of the 256 columns each
      if (grp < 2) {
        mbar_wait(y_full, (it * p.n_stages + sg) & 1);
        tc_fence_after();
        if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[10 * 64 + 8 + sg] = clock64();
        const bool last = sg + 1 == p.n_stages;
        ResidOpts ro;
        ro.no_residual = sg > 0;                 // a chained module found X / alpha in its accumulator
        if (last) {
          // every MMA of the tile has retired: the input tile (-> parameters, group 1's staging ring) and the H buffers
          // (-> group 0's ring) are dead
          resid_stage_params<D, 256>(sparam, threadIdx.x - 128, fs.b2, 0, fs.ln_mode, fs.g1, fs.be1, fs.g2, fs.be2);
          uint8_t* ring = grp == 0 ? sH : sA + kBufBytes;
          if (elected && sg == 0) resid_prefetch<D, 3, 128, 2>(ring, res_bar + grp * 4, &tmR, 0, m0, grp);
          if (p.np_blocks > 0) ro.y_smem = sH;             // y is only consumed by the projection tail: not stored
          if (p.trace && blockIdx.x == 0) ro.trace = p.trace + 11 * 64 + 32;
          ResidParams rp{nullptr, p.y_row_valid, fs.alpha, p.eps, fs.ln_mode, p.M};
          resid_ln_epilogue<D, 3, 128, 2>(tmem_y + lane_base, r, m0, 0, elected, bar_id, ring, res_bar + grp * 4, ring_phase,
                                          sparam, &tmX, &tmR, &tmY, rp, grp, 1 + kSiluGroups,
                                          reinterpret_cast<float2*>(sA + 8192), -1, -1, ro);
          if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[10 * 64 + 10 + sg] = clock64();
          if (p.np_blocks > 0) {
            // ---- projection tail (the QKV projection of the attention block that follows): y sits in sH as the A
            //      operand; each warpgroup turns 128 of a block's 256 accumulator columns into bf16 (+ bias) and
            //      stores them through two staging tiles in the dead input tile
            arrive_mma(y_ready);
            int sub_cnt = 0;
            for (int blk = 0; blk < p.np_blocks; ++blk) {
              const int ab = blk & 1;
              named_bar_sync(1 + kSiluGroups, 256);          // everybody is done with the previous block's bias
              sb1[threadIdx.x - 128] = __ldg(p.bp + blk * 256 + (threadIdx.x - 128));
              named_bar_sync(1 + kSiluGroups, 256);
              mbar_wait(q_full + ab, (it * ((p.np_blocks + 1 - ab) >> 1) + (blk >> 1)) & 1);
              tc_fence_after();
              const uint32_t tacc = (ab ? tmem_y : tmem_base) + lane_base;
#pragma unroll 1
              for (int ss = 0; ss < 2; ++ss, ++sub_cnt) {
                const int sub = grp * 2 + ss;                  // 64-column sub-tile of the block owned by this warpgroup
                uint8_t* buf = sA + (grp * 2 + (sub_cnt & 1)) * kBufBytes;
                if (elected) bulk_wait_read<1>();              // the store issued two sub-tiles ago has left this buffer
                named_bar_sync(bar_id, 128);
                uint32_t v[64];
                {
                  uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
                  uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
                  tmem_ld32(tacc + sub * 64, v0);
                  tmem_ld32(tacc + sub * 64 + 32, v1);
                }
                tmem_ld_wait();
                const float* bs = sb1 + sub * 64;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float f[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]) + bs[8 * j + e];
                  *reinterpret_cast<uint4*>(buf + sw_off(r, j)) =
                      make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (elected) { tma_store_2d(&tmP, buf, blk * 256 + sub * 64, m0); bulk_commit(); }
              }
              tc_fence_before();
              arrive_mma(q_empty + ab);
            }
            if (elected) bulk_wait_read<0>();                  // the staging tiles are free before the next tile's input lands
            named_bar_sync(bar_id, 128);
            if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[10 * 64 + 12] = clock64();
          }
          mbar_arrive(tile_done);
          if constexpr (PAIR) arrive_mma(pair_done);
          if (t + tstep < m_tiles) mbar_wait(tile_done, it & 1);   // sA / sH are re-used by the next tile
        } else {
          // first module of a chain: X and y stay on chip.  y goes straight into the input tile sA (A operand of the
          // next module), X / alpha_next stays in the accumulator columns, nothing is stored.  The residual rows arrive
          // in sA (each warpgroup's half; prefetch issued in the SiLU loop above) and are overwritten by y afterwards;
          // the parameters sit in the dead H buffers, so the weight ring keeps prefetching the next module's pieces.
          float* cparam = reinterpret_cast<float*>(sH);
          resid_stage_params<D, 256>(cparam, threadIdx.x - 128, fs.b2, 0, fs.ln_mode, fs.g1, fs.be1, fs.g2, fs.be2);
          uint8_t* ring = sA + grp * 2 * kBufBytes;
          ro.store_x = false;
          ro.y_smem = sA;
          ro.park_scale = 1.0f / p.st[sg + 1].alpha;
          if (p.trace && blockIdx.x == 0) ro.trace = p.trace + 11 * 64;
          ResidParams rp{nullptr, nullptr, fs.alpha, p.eps, fs.ln_mode, p.M};
          resid_ln_epilogue<D, 2, 128, 2>(tmem_y + lane_base, r, m0, 0, elected, bar_id, ring, res_bar + grp * 4, ring_phase,
                                          cparam, &tmX, &tmR, &tmY, rp, grp, 1 + kSiluGroups,
                                          reinterpret_cast<float2*>(sH + 8192), -1, -1, ro);
          // (the epilogue ends with a 256-thread barrier: both groups are done with the ring and the parameters)
          if (p.trace && blockIdx.x == 0 && et == 0 && grp == 0) p.trace[10 * 64 + 10 + sg] = clock64();
          arrive_mma(a_ready);                  // y in sA (fenced), X / alpha parked in Y (tcgen05.wait::st done)
        }
      }
      }
    }
    if (elected) bulk_wait_read<0>();   // the stores only have to be done READING shared memory before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  gstamp(2);
  if constexpr (CL > 1 || PAIR) cluster_sync_all();   // nobody exits while a peer may still multicast into it / signal it
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc_2sm<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

int make_2d_map(CUtensorMap* tm, bool f32, const void* base, int rows, int cols, int ld, int box_rows = 128) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)ld * (f32 ? 4 : 2)};
  const uint32_t box[2] = {(uint32_t)(f32 ? 32 : 64), (uint32_t)box_rows};
  return f32 ? tc::make_tmap_f32(tm, base, 2, dims, str, box) : tc::make_tmap_bf16(tm, base, 2, dims, str, box);
}

}  // namespace

bool ffn_fused_supported(int ld_in, int ldx, int ld_out, int M, int d, int F, int dtype, int ln_mode) {
  if (dtype != CFM_BF16 || tc::encode_tiled_fn() == nullptr) return false;
  if (d != D || F % (2 * HC) != 0 || M < 64) return false;
  if (ld_in % 8 != 0 || ldx % 4 != 0) return false;
  if (ln_mode != 0 && ld_out % 8 != 0) return false;
  return true;
}

static int ffn_launch(const void* y_in, int ld_in, const FfnModule* mods, int n, float* X, int ldx, int M, int F,
                      void* y_out, int ld_out, const uint8_t* y_row_valid, float eps, cudaStream_t st,
                      const void* Wp = nullptr, const float* bp = nullptr, void* P = nullptr, int Np = 0) {
  // cluster size along M: the CTAs of a cluster share every weight piece through TMA multicast (CFM_B200_FFN_CLUSTER)
  static const int cl_env = env_is("CFM_B200_FFN_CLUSTER", "1") ? 1 : (env_is("CFM_B200_FFN_CLUSTER", "4") ? 4 : 2);
  // CTA pairs (cta_group::2) by default: the main loop runs at 4.17 k cycles per chunk pair (4.10 k = tensor-pipe bound)
  // instead of 4.66 k, the chained tile takes 101.5 k instead of 109.3 k cycles warm (tools/ffn_chain_trace.py);
  // CFM_B200_FFN_PAIR=0 selects the single-CTA kernel (+ CFM_B200_FFN_CLUSTER multicast variants)
  static const bool pair_off = env_is("CFM_B200_FFN_PAIR", "0");
  const bool pair = !pair_off && M > BM;
  CFM_SMEM_OPT_IN((ffn_fused_kernel<1, false>), kSmemBytes);
  CFM_SMEM_OPT_IN((ffn_fused_kernel<2, false>), kSmemBytes);
  CFM_SMEM_OPT_IN((ffn_fused_kernel<4, false>), kSmemBytes);
  CFM_SMEM_OPT_IN((ffn_fused_kernel<1, true>), kSmemBytes);
  const int CL = pair ? 1 : cl_env;
  const int wbox = CL == 4 ? 64 : 128;         // rows of a weight piece one CTA fetches per TMA
  CUtensorMap tmA, tmW1[2], tmW2[2], tmX, tmY;
  int rc;
  if ((rc = make_2d_map(&tmA, false, y_in, M, D, ld_in)) != 0) return rc;
  for (int i = 0; i < n; ++i) {
    if ((rc = make_2d_map(&tmW1[i], false, mods[i].W1, F, D, D, wbox)) != 0) return rc;
    if ((rc = make_2d_map(&tmW2[i], false, mods[i].W2, D, F, F, wbox)) != 0) return rc;
  }
  if (n == 1) { tmW1[1] = tmW1[0]; tmW2[1] = tmW2[0]; }
  if ((rc = make_2d_map(&tmX, true, X, M, D, ldx)) != 0) return rc;
  tmY = tmA;
  const int ln_last = mods[n - 1].g1 ? (mods[n - 1].g2 ? 2 : 1) : 0;
  if (ln_last != 0 && Wp == nullptr && (rc = make_2d_map(&tmY, false, y_out, M, D, ld_out)) != 0) return rc;
  CUtensorMap tmWp = tmW1[0], tmP = tmA;
  if (Wp != nullptr) {
    if ((rc = make_2d_map(&tmWp, false, Wp, Np, D, D, wbox)) != 0) return rc;
    if ((rc = make_2d_map(&tmP, false, P, M, Np, Np)) != 0) return rc;
  }
  FfnParams p;
  p.bp = bp; p.np_blocks = Wp ? Np / 256 : 0;
  for (int i = 0; i < 2; ++i) {
    const FfnModule& m = mods[i < n ? i : n - 1];
    p.st[i] = FfnStage{m.b1, m.b2, m.g1, m.be1, m.g2, m.be2, m.alpha, m.g1 ? (m.g2 ? 2 : 1) : 0};
  }
  p.n_stages = n; p.y_row_valid = y_row_valid; p.eps = eps; p.M = M; p.F = F; p.trace = nullptr;
  if (const char* e = getenv("CFM_B200_FFN_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  const int m_tiles = ((M + BM - 1) / BM + CL - 1) / CL * CL;
  const int max_ctas = num_sms() / CL * CL;
  const int grid = m_tiles < max_ctas ? m_tiles : max_ctas;
  if (pair) {
    const int pair_tiles = ((M + BM - 1) / BM + 1) / 2, slots = num_sms() / 2;
    const int pgrid = 2 * (pair_tiles < slots ? pair_tiles : slots);
    CFM_CUDA_OK(launch_pdl(ffn_fused_kernel<1, true>, dim3(pgrid), dim3(kThreads), kSmemBytes, st, 2, tmA, tmW1[0], tmW2[0],
                           tmW1[1], tmW2[1], tmX, tmX, tmY, tmWp, tmP, p));
    count_variant("ffn_fused_pair");
  } else if (CL == 1)
    CFM_CUDA_OK(launch_pdl(ffn_fused_kernel<1, false>, dim3(grid), dim3(kThreads), kSmemBytes, st, 1, tmA, tmW1[0], tmW2[0], tmW1[1],
                           tmW2[1], tmX, tmX, tmY, tmWp, tmP, p));
  else if (CL == 4)
    CFM_CUDA_OK(launch_pdl(ffn_fused_kernel<4, false>, dim3(grid), dim3(kThreads), kSmemBytes, st, 4, tmA, tmW1[0], tmW2[0], tmW1[1],
                           tmW2[1], tmX, tmX, tmY, tmWp, tmP, p));
  else
    CFM_CUDA_OK(launch_pdl(ffn_fused_kernel<2, false>, dim3(grid), dim3(kThreads), kSmemBytes, st, 2, tmA, tmW1[0], tmW2[0], tmW1[1],
                           tmW2[1], tmX, tmX, tmY, tmWp, tmP, p));
  CFM_LAUNCHED_K("ffn_fused");
  return 0;
}

int ffn_fused(const void* y_in, int ld_in, const void* W1, const float* b1, const void* W2, const float* b2, float* X,
              int ldx, int M, int F, float alpha, int ln_mode, const float* g1, const float* be1, const float* g2,
              const float* be2, void* y_out, int ld_out, const uint8_t* y_row_valid, float eps, cudaStream_t st) {
  (void)ln_mode;
  const FfnModule m{W1, b1, W2, b2, alpha, g1, be1, g2, be2};
  return ffn_launch(y_in, ld_in, &m, 1, X, ldx, M, F, y_out, ld_out, y_row_valid, eps, st);
}

// Two modules on the same tiles, chained inside one kernel: the first one's X and y never leave the SM.  `a` may be
// null (single module).  With a projection (Wp != null) the last module's LayerNorm output is consumed on chip by
// P = y Wp^T + bp and not stored.
bool ffn_chain_supported(int M, int d, int F, int dtype, const FfnModule* a, const FfnModule& b, int Np) {
  if (!ffn_fused_supported(d, d, d, M, d, F, dtype, 1)) return false;
  if (a != nullptr) {
    if (a->g1 == nullptr) return false;                    // the second module's input is the first one's LayerNorm output
    // X / alpha_next is parked in the accumulator: exact only for power-of-two alpha
    int ex;
    if (!(b.alpha > 0.f) || frexpf(b.alpha, &ex) != 0.5f) return false;
  }
  if (Np != 0 && (Np % 256 != 0 || b.g1 == nullptr || b.g2 != nullptr)) return false;
  return true;
}

int ffn_chain(const void* y_in, const FfnModule* a, const FfnModule& b, float* X, int M, int F, void* y_out,
              const uint8_t* y_row_valid, float eps, const void* Wp, const float* bp, void* P, int Np, cudaStream_t st) {
  FfnModule mods[2];
  int n = 0;
  if (a != nullptr) mods[n++] = *a;
  mods[n++] = b;
  return ffn_launch(y_in, D, mods, n, X, D, M, F, y_out, D, y_row_valid, eps, st, Wp, bp, P, Np);
}

}  // namespace cfm
