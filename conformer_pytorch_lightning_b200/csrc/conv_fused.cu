// Fused Conformer convolution module for sm_100a (d = 256, kernel size 15, eval-mode BatchNorm folded):
//     X += rowmask( pw2( silu( bn( depthwise( glu( pw1( y ) ) ) ) ) ) )      [+ the LayerNorm that follows]
// replacing convolution.py:41-48 + encoder_layer.py:66-67 of the reference: one launch instead of
// GEMM(pw1+GLU) -> depthwise kernel -> GEMM(pw2+LN), and the two (tokens x d) intermediates never touch HBM.
//
// One CTA = 114 consecutive tokens of the flattened (B*T) axis: it loads the 128 rows [m0-7, m0+121) of y (7-row halo on
// each side for the 15-tap depthwise filter, re-computed by the neighbouring tile instead of exchanged), runs
//   pw1:   S_v, S_g = y W1v^T, y W1g^T        tcgen05.mma M=128 N=256 K=256, two accumulators = all 512 TMEM columns
//   GLU:   G = (S_v + b_v) * sigmoid(S_g + b_g)   8 epilogue warps, bf16 into a 32-chunk XOR-swizzled row-major smem tile
//                                                 (conflict-free for the row-wise writes AND the channel-wise reads)
//   dw:    C[o] = silu( sum_j w'[j] G[o+j] + b' ), o < 114, taps outside the token's own utterance are skipped
//          (zero padding per utterance, convolution.py:16-23); written as bf16 straight into the swizzled K-major
//          A-operand layout of the next MMA
//   pw2:   acc = C W2^T                       tcgen05.mma M=128 N=256 K=256 (re-uses S_v's TMEM columns)
//   shared residual/LayerNorm epilogue (resid_epilogue.cuh) on the 114 interior rows (TMA boxes of 114 rows).
// Weights stream through a 3 x 32 KB TMA ring ([256 rows x 64 k] pieces, 512 MMA cycles per barrier wait).
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include "resid_epilogue.cuh"

namespace cfm {
namespace {

using namespace tc;

constexpr int D = 256;
constexpr int KS = 15, HALO = 7;
constexpr int ROWS = 128 - (KS - 1);            // 114 output rows per tile
constexpr int kAtom = 16384, kPiece = 32768, NST = 3;
constexpr int kTile = 128 * D * 2;             // 64 KB
constexpr int kThreads = 512;                  // 4 control warps + 2 GLU/depthwise/epilogue warpgroups + 1 depthwise-only warpgroup
constexpr int kSmemBytes = 2 * kTile + NST * kPiece + 2 * D * 4 + 512;
static_assert(kSmemBytes <= 232448, "smem budget");

struct ConvParams {
  const float* b1;        // (512) pointwise_conv1 bias [value; gate]
  const float* dw_w;      // (15, 256) folded depthwise taps, tap-major
  const float* dw_b;      // (256) folded bias
  const float* b2;        // (256) pointwise_conv2 bias
  const float* g1; const float* be1;   // LayerNorm that follows (norm_ff), may be null
  const uint8_t* row_valid;            // pad mask on the module output (convolution.py:47-48)
  float eps;
  int M, T, ln_mode;
  long long* trace;       // optional clock64 stamps of one CTA (tools/conv_trace.py); nullptr in production
};
#define CONV_STAMP(i) do { if (p.trace && blockIdx.x == 70) p.trace[i] = clock64(); } while (0)

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {   // d += a * b on both lanes
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

// G tile: row-major [128][256] bf16, 16-byte chunk j of row r stored at chunk (j ^ (r & 31))
__device__ __forceinline__ uint32_t g_off(int r, int ch) { return r * 512 + ((((ch >> 3) ^ r) & 31) << 4) + (ch & 7) * 2; }
// C tile: A operand of pw2: 4 swizzle atoms [128 rows x 64 ch]
__device__ __forceinline__ uint32_t c_off(int o, int ch) {
  return (ch >> 6) * kAtom + o * 128 + (((((ch & 63) >> 3) ^ o) & 7) << 4) + (ch & 7) * 2;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_fused_kernel(const __grid_constant__ CUtensorMap tmYin,   // y (M,256) bf16, box 64 x 128
                  const __grid_constant__ CUtensorMap tmW1,    // (512,256) bf16, box 64 x 128
                  const __grid_constant__ CUtensorMap tmW2,    // (256,256) bf16, box 64 x 256
                  const __grid_constant__ CUtensorMap tmX,     // X (M,256) fp32, box 32 x 114 (store)
                  const __grid_constant__ CUtensorMap tmR,     // residual load (same tensor)
                  const __grid_constant__ CUtensorMap tmYout,  // y out (M,256) bf16, box 64 x 114
                  const ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sY = smem;                       // y tile (A of pw1) -> C tile (A of pw2) -> epilogue parameters + group 1's ring
  uint8_t* sG = sY + kTile;                 // G tile (written while pw1 still reads the y tile) -> group 0's staging ring
  uint8_t* sW = sG + kTile;                 // weight ring
  float* sb1 = reinterpret_cast<float*>(sW + NST * kPiece);   // [512]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb1 + 2 * D);
  uint64_t* w_full = bars;                  // [NST]
  uint64_t* w_empty = w_full + NST;         // [NST]
  uint64_t* y_full = w_empty + NST;
  uint64_t* s_full = y_full + 1;            // [2] pw1 accumulator (128 value + 128 gate channels) complete
  uint64_t* c_full = s_full + 2;            // C tile written by the 384 depthwise threads (S drained, G dead)
  uint64_t* acc_full = c_full + 1;          // pw2 accumulator complete
  uint64_t* tile_done = acc_full + 1;       // 256 arrivals
  uint64_t* res_bar = tile_done + 1;        // [2 groups][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8);
  float* sparam = reinterpret_cast<float*>(sY);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M + ROWS - 1) / ROWS;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmYin); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    prefetch_tmap(&tmX); prefetch_tmap(&tmR); prefetch_tmap(&tmYout);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
    mbar_init(y_full, 1); mbar_init(s_full, 1); mbar_init(s_full + 1, 1); mbar_init(c_full, 384); mbar_init(acc_full, 1);
    mbar_init(tile_done, 256);
    for (int s = 0; s < 8; ++s) mbar_init(res_bar + s, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0, phase = 0, it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int m0 = t * ROWS;
      if (it > 0) mbar_wait(tile_done, (it - 1) & 1);
      if (elect_one()) {
        mbar_expect_tx(y_full, kTile);
#pragma unroll
        for (int ka = 0; ka < 4; ++ka) tma_load_2d(sY + ka * kAtom, &tmYin, y_full, ka * 64, m0 - HALO);   // OOB rows -> 0
      }
      __syncwarp();
      for (int pc = 0; pc < 12; ++pc) {
        mbar_wait(w_empty + stage, phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(w_full + stage, kPiece);
          if (pc < 8) {      // pw1 piece: value rows and gate rows of 128 channels, k-slice pc & 3
            tma_load_2d(sW + stage * kPiece, &tmW1, w_full + stage, (pc & 3) * 64, (pc >> 2) * 128);
            tma_load_2d(sW + stage * kPiece + kAtom, &tmW1, w_full + stage, (pc & 3) * 64, D + (pc >> 2) * 128);
          } else {
            tma_load_2d(sW + stage * kPiece, &tmW2, w_full + stage, (pc - 8) * 64, 0);
          }
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    int stage = 0, phase = 0, it = 0;
    bool have = false;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      if (it > 0) mbar_wait(tile_done, (it - 1) & 1);
      if (lane == 0) CONV_STAMP(0);
      mbar_wait(y_full, it & 1);
      tc_fence_after();
      if (lane == 0) CONV_STAMP(1);
      const uint32_t y_addr = smem_u32(sY);
      for (int pc = 0; pc < 12; ++pc) {
        if (pc == 8) {                       // pw2 needs the C tile (and the S accumulators drained)
          if (lane == 0) CONV_STAMP(2);      // pw1 issued
          mbar_wait(c_full, it & 1);
          tc_fence_after();
          if (lane == 0) CONV_STAMP(3);
        }
        if (!have) mbar_wait(w_full + stage, phase);
        tc_fence_after();
        {
          const int ns = (stage + 1 == NST) ? 0 : stage + 1;
          have = mbar_test(w_full + ns, (stage + 1 == NST) ? (phase ^ 1) : phase);
        }
        if (elect_one()) {
          const int kc = pc & 3;
          const uint64_t da = umma_desc_sw128(y_addr + kc * kAtom);
          const uint64_t db = umma_desc_sw128(smem_u32(sW + stage * kPiece));
          const uint32_t d = tmem_base + ((pc >= 4 && pc < 8) ? 256 : 0);   // pw1 accumulator 1 / pw1 acc 0 and pw2
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0);
          umma_commit(w_empty + stage);
          if (pc == 3) umma_commit(s_full);         // channels [0,128): GLU starts while the second half runs
          if (pc == 7) umma_commit(s_full + 1);
          if (pc == 11) umma_commit(acc_full);
        }
        __syncwarp();
        if (pc == 11 && lane == 0) CONV_STAMP(4);
        if (++stage == NST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (3 warpgroups) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                  // tile row for the row-wise phases
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - grp * 128;
    const int tid = threadIdx.x - 128;            // 0..383
    const bool elected = (et == 0);
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    uint32_t ring_phase = 0;
    // depthwise role: channel pair + third of the output rows (warpgroup 2 only takes part in the depthwise phase)
    const int cp = tid & 127, third = tid >> 7;
    float2 wt[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(p.dw_w + j * D + 2 * cp));
    const float2 wb = __ldg(reinterpret_cast<const float2*>(p.dw_b + 2 * cp));
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int m0 = t * ROWS;
      if (grp < 2) {
      sb1[tid] = p.b1[tid];
      sb1[256 + tid] = p.b1[256 + tid];
      named_bar_sync(3, 256);
      // ---- GLU: accumulator hh = [value | gate] of channels [128 hh, 128 hh + 128) -> G (bf16, swizzled row-major);
      //      each warpgroup takes 64 of the channels
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        if (tid == 0) CONV_STAMP(8 + hh);
        mbar_wait(s_full + hh, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int ch0 = hh * 128 + grp * 64 + cc * 32;
          const uint32_t col = hh * 256 + grp * 64 + cc * 32;
          uint32_t v[32], g[32];
          tmem_ld32(tmem_base + lane_base + col, v);
          tmem_ld32(tmem_base + lane_base + col + 128, g);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = ch0 + 8 * j + e;
              f[e] = (__uint_as_float(v[8 * j + e]) + sb1[c]) * sigmoid_fast(__uint_as_float(g[8 * j + e]) + sb1[256 + c]);
            }
            *reinterpret_cast<uint4*>(sG + g_off(r, ch0 + 8 * j)) =
                make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          }
        }
      }
      tc_fence_before();
      if (tid == 0) CONV_STAMP(10);
      }
      named_bar_sync(5, 384);                       // G complete; pw1 has retired, so the y tile and S accumulators are dead
      if (tid == 0) CONV_STAMP(11);
      // ---- depthwise conv + folded BatchNorm + SiLU -> C (A operand of pw2)
      constexpr int PASS = 19;                      // 2 passes x 19 outputs = 38 rows per thread
#pragma unroll 1
      for (int ps = 0; ps < 2; ++ps) {
        const int o0 = third * 38 + ps * PASS;
        const int Ro0 = m0 + o0;                    // global token of the first output of the pass
        // utterance of each output token: taps reaching outside [lo, hi) are zero padding
        const int u0 = Ro0 / p.T, u1 = (Ro0 + PASS - 1) / p.T;
        const bool interior = (u0 == u1) && (Ro0 - HALO >= u0 * p.T) && (Ro0 + PASS - 1 + HALO < (u0 + 1) * p.T) &&
                              (Ro0 + PASS - 1 < p.M);
        float2 acc[PASS];
#pragma unroll
        for (int i = 0; i < PASS; ++i) acc[i] = wb;
        if (interior) {
#pragma unroll
          for (int s = 0; s < PASS + KS - 1; ++s) {
            const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sG + g_off(o0 + s, 2 * cp)));
#pragma unroll
            for (int i = 0; i < PASS; ++i) {
              const int j = s - i;
              if (j >= 0 && j < KS) ffma2(acc[i], xv, wt[j]);
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < PASS + KS - 1; ++s) {
            const int Rin = Ro0 - HALO + s;          // global token of this input row
            const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(sG + g_off(o0 + s, 2 * cp)));
#pragma unroll
            for (int i = 0; i < PASS; ++i) {
              const int j = s - i;
              if (j >= 0 && j < KS) {
                const int Ro = Ro0 + i;
                const int lo = (Ro / p.T) * p.T;
                const bool ok = (Ro < p.M) && (Rin >= lo) && (Rin < lo + p.T);
                if (ok) ffma2(acc[i], xv, wt[j]);
              }
            }
          }
        }
#pragma unroll
        for (int i = 0; i < PASS; ++i)
          *reinterpret_cast<uint32_t*>(sY + c_off(o0 + i, 2 * cp)) = pack_bf16x2(silu_fast(acc[i].x), silu_fast(acc[i].y));
      }
      fence_proxy_async_smem();
      mbar_arrive(c_full);
      if (tid == 0) CONV_STAMP(12);
      if (grp == 2) continue;                       // (its next stop is the G-complete barrier of the next tile)
      // ---- pw2 accumulator -> residual stream (+ LayerNorm), interior rows only; each warpgroup takes 128 columns
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      if (tid == 0) CONV_STAMP(13);
      // every MMA of the tile has retired: C (-> parameters, group 1's staging ring) and G (-> group 0's ring) are dead
      resid_stage_params<D, 256>(sparam, tid, p.b2, 0, p.ln_mode, p.g1, p.be1, nullptr, nullptr);
      uint8_t* ring = grp == 0 ? sG : sY + kBufBytes;
      if (elected) resid_prefetch<D, 3, ROWS, 2>(ring, res_bar + grp * 4, &tmR, 0, m0, grp);
      ResidParams rp{p.row_valid, nullptr, 1.0f, p.eps, p.ln_mode, p.M};
      resid_ln_epilogue<D, 3, ROWS, 2>(tmem_base + lane_base, r, m0, 0, elected, 1 + grp, ring, res_bar + grp * 4, ring_phase,
                                       sparam, &tmX, &tmR, &tmYout, rp, grp, 4, reinterpret_cast<float2*>(sY + 8192));
      mbar_arrive(tile_done);
      if (tid == 0) CONV_STAMP(14);
      if (t + (int)gridDim.x < n_tiles) mbar_wait(tile_done, it & 1);   // sb1 / sY / sG are re-used by the next tile
    }
    if (elected && grp < 2) bulk_wait_read<0>();   // the stores only have to be done READING shared memory before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int make_map(CUtensorMap* tm, bool f32, const void* base, int rows, int cols, int ld, int box_rows) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t str[1] = {(uint64_t)ld * (f32 ? 4 : 2)};
  const uint32_t box[2] = {(uint32_t)(f32 ? 32 : 64), (uint32_t)box_rows};
  return f32 ? tc::make_tmap_f32(tm, base, 2, dims, str, box) : tc::make_tmap_bf16(tm, base, 2, dims, str, box);
}

}  // namespace

bool conv_fused_supported(int M, int T, int d, int k, int dtype) {
  return dtype == CFM_BF16 && tc::encode_tiled_fn() != nullptr && d == D && k == KS && M >= 64 && T >= KS;
}

int conv_fused(const void* y_in, const void* W1, const float* b1, const float* dw_w, const float* dw_b, const void* W2,
               const float* b2, float* X, int M, int T, const uint8_t* row_valid, const float* g1, const float* be1,
               void* y_out, float eps, cudaStream_t st) {
  CFM_SMEM_OPT_IN(conv_fused_kernel, kSmemBytes);
  CUtensorMap tmYin, tmW1, tmW2, tmX, tmYout;
  int rc;
  if ((rc = make_map(&tmYin, false, y_in, M, D, D, 128)) != 0) return rc;
  if ((rc = make_map(&tmW1, false, W1, 2 * D, D, D, 128)) != 0) return rc;
  if ((rc = make_map(&tmW2, false, W2, D, D, D, 256)) != 0) return rc;
  if ((rc = make_map(&tmX, true, X, M, D, D, ROWS)) != 0) return rc;
  tmYout = tmYin;
  if (g1 != nullptr && (rc = make_map(&tmYout, false, y_out, M, D, D, ROWS)) != 0) return rc;
  ConvParams p{b1, dw_w, dw_b, b2, g1, be1, row_valid, eps, M, T, g1 ? 1 : 0, nullptr};
  if (const char* e = getenv("CFM_B200_CONV_TRACE_PTR")) p.trace = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  const int n_tiles = (M + ROWS - 1) / ROWS;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  CFM_CUDA_OK(launch_pdl(conv_fused_kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, 1, tmYin, tmW1, tmW2, tmX, tmX, tmYout, p));
  CFM_LAUNCHED_K("conv_fused");
  return 0;
}

}  // namespace cfm
