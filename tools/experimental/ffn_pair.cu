// EXPERIMENT, not part of libcfm_b200.so (removed from the product build in round 2): cta_group::2 variant of the fused
// feed-forward kernel.  Numerically correct, 45.9 us vs 30.8 us for ffn_fused.cu; kept as the starting point for the
// 'two tiles in flight per SM' lead of DESIGN.md section 8.  Builds against csrc/*.cuh.
// Fused macaron feed-forward, CTA-pair edition (tcgen05 cta_group::2) for sm_100a, d = 256.
//
// Same algorithm and schedule as ffn_fused.cu (see there), but every tcgen05.mma spans the two SMs of a 2-CTA
// cluster: M = 256 (128 token rows per CTA), and each CTA holds only HALF of every weight piece in its shared
// memory (the tensor cores read B rows from both SMs).  ffn_fused.cu is shared-memory-bandwidth bound (measured:
// operand reads 128 B/clk for N = 128 MMAs + 64 B/clk of TMA weight fill + H stores > the 128 B/clk an SM has);
// the pair halves both the weight fill and the B-operand reads per SM:
//     per 128-unit hidden chunk and SM:  G1 96 KB + G2 64 KB + TMA 64 KB + H 32 KB = 256 KB  (was 384 KB)
// Roles per CTA: warp 0 TMA producer (own A tile, own half of each weight piece; completion bytes are credited to the
// LEADER's mbarriers), warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator (cta_group::2, both CTAs),
// warps 4-11 SiLU / residual+LayerNorm epilogue on the CTA's own 128 rows.  Cross-CTA signalling:
//   leader <- peer : s_empty / h_full / pair_done arrivals (mbarrier.arrive.release.cluster on mapa'd addresses)
//   leader -> both : tcgen05.commit.cta_group::2 ... multicast::cluster (w_empty, s_full, h_empty, y_full)
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include "resid_epilogue.cuh"
#include <stdlib.h>

namespace cfm {
namespace {

using namespace tc;

constexpr int D = 256, HC = 128, BM = 128;
constexpr int kAtom = 16384;        // 128 rows x 64 k bf16
constexpr int kSlot = 16384;        // this CTA's half of one weight piece
constexpr int NST = 5;
constexpr int kABytes = BM * D * 2;
constexpr int kHBytes = BM * HC * 2;
constexpr int kThreads = 384;
constexpr int kSmemBytes = kABytes + 2 * kHBytes + NST * kSlot + 2 * HC * 4 + 512;
static_assert(kSmemBytes <= 232448, "smem budget");

struct PairParams {
  const float* b1; const float* b2;
  const float* g1; const float* be1; const float* g2; const float* be2;
  const uint8_t* y_row_valid;
  float alpha, eps;
  int M, F, ln_mode;
};

__device__ __forceinline__ void job_of(int jx, int NC, bool& g1, int& c) {
  if (jx < 2) { g1 = true; c = jx; }
  else if (jx >= 2 * NC - 2) { g1 = false; c = jx - NC; }
  else if (jx & 1) { g1 = true; c = (jx + 1) >> 1; }
  else { g1 = false; c = (jx - 2) >> 1; }
}

__global__ void __launch_bounds__(kThreads, 1)
ffn_pair_kernel(const __grid_constant__ CUtensorMap tmA,    // y  (M, 256) bf16, box 64 x 128
                const __grid_constant__ CUtensorMap tmW1,   // W1 (F, 256) bf16, box 64 x 64   (half of a 128-row chunk)
                const __grid_constant__ CUtensorMap tmW2,   // W2 (256, F) bf16, box 64 x 128  (half of the 256 rows)
                const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmY, const PairParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sH = sA + kABytes;
  uint8_t* sW = sH + 2 * kHBytes;
  float* sparam = reinterpret_cast<float*>(sA);               // aliases the dead input tile in the final epilogue
  float* sb1 = reinterpret_cast<float*>(sW + NST * kSlot);    // [2][HC]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb1 + 2 * HC);
  uint64_t* w_full = bars;                 // [NST] leader's instance is used (bytes of both CTAs)
  uint64_t* w_empty = w_full + NST;        // [NST] both CTAs (multicast commit)
  uint64_t* a_full = w_empty + NST;        // leader's instance (both CTAs' A tiles)
  uint64_t* s_full = a_full + 1;           // [2] both CTAs (multicast commit)
  uint64_t* s_empty = s_full + 2;          // [2] leader's instance, 512 arrivals
  uint64_t* h_full = s_empty + 2;          // [2] leader's instance, 512 arrivals
  uint64_t* h_empty = h_full + 2;          // [2] both CTAs (multicast commit)
  uint64_t* y_full = h_empty + 2;          // both CTAs (multicast commit)
  uint64_t* tile_done = y_full + 1;        // local: this CTA's final epilogue finished (128 arrivals)
  uint64_t* pair_done = tile_done + 1;     // leader's instance: both CTAs' final epilogues finished (256 arrivals)
  uint64_t* res_bar = pair_done + 1;       // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = (crank == 0);
  const int m_tiles = ((p.M + BM - 1) / BM + 1) / 2 * 2;      // phantom tile past M keeps the pair in lock-step
  const int NC = p.F / HC;
  const int n_jobs = 2 * NC;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    prefetch_tmap(&tmX); prefetch_tmap(&tmR); prefetch_tmap(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
    mbar_init(a_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + s, 1); mbar_init(s_empty + s, 512);
      mbar_init(h_full + s, 512); mbar_init(h_empty + s, 1);
    }
    mbar_init(y_full, 1);
    mbar_init(tile_done, 128);
    mbar_init(pair_done, 256);
    for (int s = 0; s < 4; ++s) mbar_init(res_bar + s, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers initialised, TMEM allocated in both SMs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_y = tmem_base + 256;

  // shared::cluster addresses of the leader's barriers that this CTA signals
  const uint32_t a_full_ldr = mapa_u32(smem_u32(a_full), 0);
  const uint32_t w_full_ldr = mapa_u32(smem_u32(w_full), 0);
  const uint32_t s_empty_ldr = mapa_u32(smem_u32(s_empty), 0);
  const uint32_t h_full_ldr = mapa_u32(smem_u32(h_full), 0);
  const uint32_t pair_done_ldr = mapa_u32(smem_u32(pair_done), 0);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      const int m0 = t * BM;
      if (it > 0) mbar_wait(tile_done, (it - 1) & 1);
      if (elect_one()) {
        if (leader) mbar_expect_tx(a_full, 2 * kABytes);
#pragma unroll
        for (int ka = 0; ka < D / 64; ++ka) tma_load_2d_2sm(sA + ka * kAtom, &tmA, a_full_ldr, ka * 64, m0);
      }
      __syncwarp();
      for (int jx = 0; jx < n_jobs; ++jx) {
        bool g1; int c;
        job_of(jx, NC, g1, c);
        for (int pc = 0; pc < 2; ++pc) {
          mbar_wait(w_empty + stage, phase ^ 1);
          if (elect_one()) {
            uint8_t* dst = sW + stage * kSlot;
            if (leader) mbar_expect_tx(w_full + stage, 2 * kSlot);
            const uint32_t bar = w_full_ldr + stage * 8;
            if (g1) {     // my 64 of the 128 hidden rows of chunk c, k in [pc*128, +128): two 8 KB atoms
              tma_load_2d_2sm(dst, &tmW1, bar, pc * 128, c * HC + crank * 64);
              tma_load_2d_2sm(dst + kSlot / 2, &tmW1, bar, pc * 128 + 64, c * HC + crank * 64);
            } else {      // my 128 of the 256 output rows, hidden k in [c*128 + pc*64, +64)
              tma_load_2d_2sm(dst, &tmW2, bar, c * HC + pc * 64, crank * 128);
            }
          }
          __syncwarp();
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA; converged warp, one elected lane) =====================
    constexpr uint32_t idesc1 = umma_idesc_bf16(256, 128);   // G1: M = 256 (pair), N = 128 hidden units
    constexpr uint32_t idesc2 = umma_idesc_bf16(256, 256);   // G2: M = 256 (pair), N = 256 outputs
    constexpr uint16_t kBoth = 0x3;
    int stage = 0, phase = 0, it = 0;
    uint32_t n_se0 = 0, n_se1 = 0, n_hf0 = 0, n_hf1 = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      if (it > 0) mbar_wait_cluster(pair_done, (it - 1) & 1);   // both CTAs drained Y / S of the previous tile pair
      mbar_wait(a_full, it & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), h_addr = smem_u32(sH);
      for (int jx = 0; jx < n_jobs; ++jx) {
        bool g1; int c;
        job_of(jx, NC, g1, c);
        const int b = c & 1;
        if (g1) {
          uint32_t& n_se = b ? n_se1 : n_se0;
          mbar_wait_cluster(s_empty + b, (n_se & 1) ^ 1);
          ++n_se;
          tc_fence_after();
          for (int pc = 0; pc < 2; ++pc) {
            mbar_wait(w_full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t w_addr = smem_u32(sW + stage * kSlot);
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                const uint64_t da = umma_desc_sw128(a_addr + (2 * pc + a) * kAtom);
                const uint64_t db = umma_desc_sw128(w_addr + a * (kSlot / 2));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_2sm(tmem_base + b * HC, da + 2 * k, db + 2 * k, idesc1, (pc | a | k) != 0);
              }
              umma_commit_2sm(w_empty + stage, kBoth);
              if (pc == 1) umma_commit_2sm(s_full + b, kBoth);
            }
            __syncwarp();
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        } else {
          uint32_t& n_hf = b ? n_hf1 : n_hf0;
          mbar_wait_cluster(h_full + b, n_hf & 1);
          ++n_hf;
          tc_fence_after();
          for (int pc = 0; pc < 2; ++pc) {
            mbar_wait(w_full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = umma_desc_sw128(h_addr + b * kHBytes + pc * kAtom);
              const uint64_t db = umma_desc_sw128(smem_u32(sW + stage * kSlot));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_2sm(tmem_y, da + 2 * k, db + 2 * k, idesc2, (c | pc | k) != 0);
              umma_commit_2sm(w_empty + stage, kBoth);
              if (pc == 1) {
                umma_commit_2sm(h_empty + b, kBoth);
                if (jx == n_jobs - 1) umma_commit_2sm(y_full, kBoth);
              }
            }
            __syncwarp();
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - grp * 128;
    const bool elected = (et == 0);
    const int bar_id = 1 + grp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint32_t ring_phase = 0;
    uint32_t n_sf0 = 0, n_sf1 = 0, n_he0 = 0, n_he1 = 0;
    int it = 0;
    for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
      const int m0 = t * BM;
      if (it > 0 && grp == 1) mbar_wait(tile_done, (it - 1) & 1);
      for (int c = 0; c < NC; ++c) {
        const int b = c & 1;
        if (et < 64) sb1[b * HC + grp * 64 + et] = p.b1[c * HC + grp * 64 + et];
        named_bar_sync(bar_id, 128);
        uint32_t& n_sf = b ? n_sf1 : n_sf0;
        mbar_wait(s_full + b, n_sf & 1);
        ++n_sf;
        tc_fence_after();
        uint32_t v[64];
        {
          uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          tmem_ld32(tmem_base + lane_base + b * HC + grp * 64, v0);
          tmem_ld32(tmem_base + lane_base + b * HC + grp * 64 + 32, v1);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_cluster(s_empty_ldr + b * 8);          // leader collects both CTAs' drains
        const float* bs = sb1 + b * HC + grp * 64;
        uint4 pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = silu_fast(__uint_as_float(v[8 * j + e]) + bs[8 * j + e]);
          pk[j] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        uint32_t& n_he = b ? n_he1 : n_he0;
        mbar_wait(h_empty + b, (n_he & 1) ^ 1);
        ++n_he;
        uint8_t* hb = sH + b * kHBytes + grp * kAtom;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(hb + sw_off(r, j)) = pk[j];
        fence_proxy_async_smem();
        mbar_arrive_cluster(h_full_ldr + b * 8);
      }
      if (grp == 0) {
        mbar_wait(y_full, it & 1);
        tc_fence_after();
        resid_stage_params<D>(sparam, et, p.b2, 0, p.ln_mode, p.g1, p.be1, p.g2, p.be2);
        if (elected) resid_prefetch<D, 4>(sH, res_bar, &tmR, 0, m0);
        ResidParams rp{nullptr, p.y_row_valid, p.alpha, p.eps, p.ln_mode, p.M};
        resid_ln_epilogue<D, 4>(tmem_y + lane_base, r, m0, 0, elected, bar_id, sH, res_bar, ring_phase, sparam, &tmX,
                                &tmR, &tmY, rp);
        mbar_arrive(tile_done);
        mbar_arrive_cluster(pair_done_ldr);
      }
    }
    if (elected) bulk_wait_read<0>();   // the stores only have to be done READING shared memory before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // nobody frees TMEM / exits while the peer may still touch it
  if (warp == 2) tmem_dealloc_2sm<512>(tmem_base);
}

int make_map(CUtensorMap* tm, bool f32, const void* base, int rows, int cols, int ld, int box_rows) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)ld * (f32 ? 4 : 2)};
  const uint32_t box[2] = {(uint32_t)(f32 ? 32 : 64), (uint32_t)box_rows};
  return f32 ? tc::make_tmap_f32(tm, base, 2, dims, str, box) : tc::make_tmap_bf16(tm, base, 2, dims, str, box);
}

}  // namespace

int ffn_pair(const void* y_in, int ld_in, const void* W1, const float* b1, const void* W2, const float* b2, float* X,
             int ldx, int M, int F, float alpha, int ln_mode, const float* g1, const float* be1, const float* g2,
             const float* be2, void* y_out, int ld_out, const uint8_t* y_row_valid, float eps, cudaStream_t st) {
  CFM_SMEM_OPT_IN(ffn_pair_kernel, kSmemBytes);
  CUtensorMap tmA, tmW1, tmW2, tmX, tmY;
  int rc;
  if ((rc = make_map(&tmA, false, y_in, M, D, ld_in, 128)) != 0) return rc;
  if ((rc = make_map(&tmW1, false, W1, F, D, D, 64)) != 0) return rc;
  if ((rc = make_map(&tmW2, false, W2, D, F, F, 128)) != 0) return rc;
  if ((rc = make_map(&tmX, true, X, M, D, ldx, 128)) != 0) return rc;
  tmY = tmA;
  if (ln_mode != 0 && (rc = make_map(&tmY, false, y_out, M, D, ld_out, 128)) != 0) return rc;
  PairParams p{b1, b2, g1, be1, g2, be2, y_row_valid, alpha, eps, M, F, ln_mode};
  const int m_tiles = ((M + BM - 1) / BM + 1) / 2 * 2;
  const int max_ctas = num_sms() / 2 * 2;
  const int grid = m_tiles < max_ctas ? m_tiles : max_ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CFM_CUDA_OK(cudaLaunchKernelEx(&cfg, ffn_pair_kernel, tmA, tmW1, tmW2, tmX, tmX, tmY, p));
  CFM_LAUNCHED_K("ffn_pair");
  return 0;
}

}  // namespace cfm
