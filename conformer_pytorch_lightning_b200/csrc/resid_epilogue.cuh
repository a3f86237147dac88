// Residual-stream epilogue shared by the tcgen05 GEMM (gemm_tc.cu) and the fused FFN kernel (ffn_fused.cu):
//   v = R + alpha * rowmask(acc + bias)                    (encoder_layer.py:58,62,66,69 of the reference)
//   ln_mode 0:  X = v
//   ln_mode 1:  X = v,          y = ymask(LN(v; g1,b1))    (encoder_layer.py:59,63,67)
//   ln_mode 2:  X = LN(v;g1,b1), y = ymask(LN(X; g2,b2))   (encoder_layer.py:70 chained with :56 of the next layer)
// Thread = output row: tcgen05.ld 32x32b gives each thread 32 consecutive fp32 columns of ITS row, so LayerNorm
// statistics need no shuffle reduction.  NG = 1: one warpgroup (128 threads) walks all BN columns.  NG = 2: two
// warpgroups split the columns of every row in halves (a single warp per scheduler cannot hide the tcgen05.ld /
// shared-memory latencies of this epilogue, two can) and exchange their partial row sums through shared memory; each
// group owns a staging ring of R tiles, an elected thread and R mbarriers.  The fp32 residual arrives through TMA loads into a ring of
// 128-byte-swizzled [128 x 128 B] staging tiles, X and y leave through TMA stores from the same ring, and between
// the passes the pre-norm row is parked in the tile's own TMEM accumulator columns (tcgen05.st).
#pragma once
#include "cfm_common.cuh"
#include "tc_common.cuh"

namespace cfm {
namespace tc {

constexpr int kBufBytes = 128 * 128;   // one staging tile: 128 rows x 128 bytes (64 bf16 or 32 fp32 columns)

// byte offset of 16-byte chunk j of row r inside a 128B-swizzled [128 x 128 B] tile
__device__ __forceinline__ uint32_t sw_off(int r, int j) { return r * 128 + (((j ^ r) & 7) << 4); }

// 2-D tensor maps address rows of the flattened (tokens, N) matrix; with z >= 0 the maps are 3-D (N, T, B) and the tile
// is rows [m0, m0+ROWS) of utterance z, so that rows past the utterance's end are clipped instead of running into the
// next utterance.
__device__ __forceinline__ void resid_tma_load(void* dst, const CUtensorMap* tm, uint64_t* bar, int c, int m0, int z) {
  if (z < 0) tma_load_2d(dst, tm, bar, c, m0); else tma_load_3d(dst, tm, bar, c, m0, z);
}
__device__ __forceinline__ void resid_tma_store(const CUtensorMap* tm, const void* src, int c, int m0, int z) {
  if (z < 0) tma_store_2d(tm, src, c, m0); else tma_store_3d(tm, src, c, m0, z);
}

// Variations used when two feed-forward modules are chained inside one kernel (ffn_fused.cu):
struct ResidOpts {
  bool store_x = true;          // false: X is not written to global memory (it is consumed on chip)
  uint8_t* y_smem = nullptr;    // != nullptr: y is written here as four [128 x 64] bf16 A-operand atoms instead of stored
  float park_scale = 1.f;       // the X row left in the TMEM accumulator columns is multiplied by this (it becomes the
                                // initial value of the next module's accumulator: X / alpha_next)
  bool no_residual = false;     // the accumulator already contains the residual: v = alpha * (acc + bias)
  long long* trace = nullptr;   // optional clock64 stamps of (group 0, row 0): tools/mhsa_trace.py
};

struct ResidParams {
  const uint8_t* row_valid;     // rows whose GEMM result is forced to 0 (pad mask)
  const uint8_t* y_row_valid;   // rows of y forced to 0
  float alpha, eps;
  int ln_mode;
  int M;
};

// sparam layout: [0,BN) bias  [BN,2BN) g1  [2BN,3BN) b1  [3BN,4BN) g2  [4BN,5BN) b2
template <int BN, int NTHREADS = 128>
__device__ __forceinline__ void resid_stage_params(float* sparam, int et, const float* bias, int n0, int ln_mode,
                                                   const float* g1, const float* b1, const float* g2, const float* b2) {
  for (int i = et; i < BN; i += NTHREADS) {
    sparam[i] = bias ? bias[n0 + i] : 0.f;
    if (ln_mode >= 1) { sparam[BN + i] = g1[i]; sparam[2 * BN + i] = b1[i]; }
    if (ln_mode == 2) { sparam[3 * BN + i] = g2[i]; sparam[4 * BN + i] = b2[i]; }
  }
}

// issue the TMA loads of the first min(R, BN/32) residual chunks (elected thread only).  ROWS = rows of the TMA box
// (128 for GEMM tiles; the fused conv module stores only its 128-(k-1) interior rows).
// (ring, res_bar: the calling group's own; grp selects the group's column half when NG = 2)
template <int BN, int R, int ROWS = 128, int NG = 1>
__device__ __forceinline__ void resid_prefetch(uint8_t* ring, uint64_t* res_bar, const CUtensorMap* tmR, int n0, int m0,
                                               int grp = 0, int z = -1) {
  constexpr int NCHG = BN / 32 / NG;
#pragma unroll
  for (int c = 0; c < (NCHG < R ? NCHG : R); ++c) {
    mbar_expect_tx(res_bar + c, ROWS * 128);
    resid_tma_load(ring + c * kBufBytes, tmR, res_bar + c, n0 + (grp * NCHG + c) * 32, m0, z);
  }
}

// Body of the epilogue.  Preconditions: accumulator complete and visible (tfull waited + tcgen05 after-sync fence),
// resid_stage_params + resid_prefetch done.  `taddr` = TMEM address of this thread's row, column 0 of the accumulator.
// `bar_id` names a 128-thread barrier private to the calling warpgroup.  On return every TMEM access of the thread
// has completed and all TMA stores have finished READING the ring (it may be overwritten).
// NG = 2: `grp` = 0/1, `xbar` names a 256-thread barrier shared by both groups, `xch` = 512 float2 of shared memory;
// sparam must have been staged by the time every thread passes the first barrier (256-thread barrier here).
template <int BN, int R, int ROWS = 128, int NG = 1>
__device__ __forceinline__ void resid_ln_epilogue(uint32_t taddr, int r, int m0, int n0, bool elected, int bar_id,
                                                  uint8_t* ring, uint64_t* res_bar, uint32_t& ring_phase,
                                                  const float* sparam, const CUtensorMap* tmX, const CUtensorMap* tmR,
                                                  const CUtensorMap* tmY, const ResidParams& p, int grp = 0, int xbar = 0,
                                                  float2* xch = nullptr, int z = -1, int grow0 = -1,
                                                  const ResidOpts o = ResidOpts()) {
  constexpr int NCH = BN / 32 / NG;              // 32-column fp32 chunks per row handled by this group
  const int ch0 = grp * NCH;                     // first chunk of this group
  const int ln = p.ln_mode;
  const int row = (grow0 >= 0 ? grow0 : m0) + r;   // row of the flattened (tokens, N) matrix: mask lookups
  const bool row_ok = row < p.M;
  const bool valid = (p.row_valid == nullptr) || !row_ok || (p.row_valid[row] != 0);
  const float a = valid ? p.alpha : 0.f;
  if (NG == 1) named_bar_sync(bar_id, 128); else named_bar_sync(xbar, 128 * NG);   // sparam visible
#define RTR(i) do { if (o.trace && grp == 0 && r == 0) o.trace[i] = clock64(); } while (0)
  RTR(0);
  float s1 = 0.f, s2 = 0.f;
  // ---- pass 1: v = R + alpha*(acc + bias)
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    const int b = c % R;
    uint8_t* buf = ring + b * kBufBytes;
    if (!o.no_residual) {
      mbar_wait(res_bar + b, (ring_phase >> b) & 1u);
      ring_phase ^= (1u << b);
    } else if (c >= R) {                         // no load orders the re-use of the buffer: wait for its last store
      if (elected) bulk_wait_read<R - 1>();
      named_bar_sync(bar_id, 128);
    }
    RTR(1 + 2 * c);
    uint32_t v[32];
    tmem_ld32(taddr + (ch0 + c) * 32, v);
    tmem_ld_wait();
    if (c == 1) RTR(13);
    // all shared-memory reads of the chunk first (the residual cells and the bias as 16-byte loads), the arithmetic on
    // registers, then the writes: a load -> use -> store chain per 16-byte cell serialises on the shared-memory latency
    const float4* bs4 = reinterpret_cast<const float4*>(sparam + (ch0 + c) * 32);
    float4 xr[8];
    if (!o.no_residual) {
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[j] = *reinterpret_cast<const float4*>(buf + sw_off(r, j));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) xr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ln 1: what stays parked is X itself (scaled for a chained module); ln 2: the pre-norm row, unscaled
    const float ps = ln == 1 ? o.park_scale : 1.f;
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;      // two independent accumulation chains
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 bq = bs4[j];
      float4 x = xr[j];
      x.x = fmaf(a, __uint_as_float(v[4 * j]) + bq.x, x.x);
      x.y = fmaf(a, __uint_as_float(v[4 * j + 1]) + bq.y, x.y);
      x.z = fmaf(a, __uint_as_float(v[4 * j + 2]) + bq.z, x.z);
      x.w = fmaf(a, __uint_as_float(v[4 * j + 3]) + bq.w, x.w);
      if (j & 1) { s1b += (x.x + x.y) + (x.z + x.w); s2b += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w); }
      else { s1a += (x.x + x.y) + (x.z + x.w); s2a += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w); }
      v[4 * j] = __float_as_uint(x.x * ps); v[4 * j + 1] = __float_as_uint(x.y * ps);
      v[4 * j + 2] = __float_as_uint(x.z * ps); v[4 * j + 3] = __float_as_uint(x.w * ps);
      xr[j] = x;
    }
    s1 += s1a + s1b;
    s2 += s2a + s2b;
    if (ln != 2 && o.store_x) {                  // X chunk leaves through the same buffer
#pragma unroll
      for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(buf + sw_off(r, j)) = xr[j];
    }
    if (c == 1) RTR(14);
    if (ln != 0) tmem_st32(taddr + (ch0 + c) * 32, v);   // park the row in our accumulator columns
    if (c == 1) RTR(15);
    fence_proxy_async_smem();
    RTR(2 + 2 * c);
    named_bar_sync(bar_id, 128);
    if (c == 1) RTR(16);
    if (elected) {
      if (ln != 2 && o.store_x) {
        resid_tma_store(tmX, buf, n0 + (ch0 + c) * 32, m0, z);
        bulk_commit();
        // refill the PREVIOUS chunk's buffer once its store has finished reading it
        if (!o.no_residual && c >= 1 && c - 1 + R < NCH) {
          bulk_wait_read<1>();
          const int pb = (c - 1) % R;
          mbar_expect_tx(res_bar + pb, ROWS * 128);
          resid_tma_load(ring + pb * kBufBytes, tmR, res_bar + pb, n0 + (ch0 + c - 1 + R) * 32, m0, z);
        }
      } else if (!o.no_residual && c + R < NCH) {   // nothing is stored in pass 1: buffer b is free right away
        mbar_expect_tx(res_bar + b, ROWS * 128);
        resid_tma_load(buf, tmR, res_bar + b, n0 + (ch0 + c + R) * 32, m0, z);
      }
    }
  }
  RTR(9);
  if (ln != 0) {
    tmem_st_wait();
    if (NG == 2) {                               // row sums of the other half of the columns
      xch[grp * 128 + r] = make_float2(s1, s2);
      named_bar_sync(xbar, 256);
      const float2 o = xch[(grp ^ 1) * 128 + r];
      s1 += o.x; s2 += o.y;
    }
    const float inv_n = 1.0f / BN;
    float mean = s1 * inv_n;
    float rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + p.eps);
    if (elected) bulk_wait_read<0>();
    // every staging buffer is free again (of BOTH groups when y goes to a shared-memory tile that may overlap them)
    if (NG == 2 && o.y_smem) named_bar_sync(xbar, 256); else named_bar_sync(bar_id, 128);
    int nbuf = 0;
    if (ln == 2) {
      // ---- pass 2 (double LN): X = LN1(v) -> TMEM + fp32 TMA store, statistics of X
      s1 = 0.f; s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c, ++nbuf) {
        uint8_t* buf = ring + (nbuf % R) * kBufBytes;
        if (o.store_x && nbuf >= R) { if (elected) bulk_wait_read<R - 1>(); named_bar_sync(bar_id, 128); }
        uint32_t v[32];
        tmem_ld32(taddr + (ch0 + c) * 32, v);
        tmem_ld_wait();
        const float4* g4 = reinterpret_cast<const float4*>(sparam + BN + (ch0 + c) * 32);
        const float4* be4 = reinterpret_cast<const float4*>(sparam + 2 * BN + (ch0 + c) * 32);
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 gq = g4[j], bq = be4[j];
          float4 x;
          x.x = fmaf((__uint_as_float(v[4 * j]) - mean) * rstd, gq.x, bq.x);
          x.y = fmaf((__uint_as_float(v[4 * j + 1]) - mean) * rstd, gq.y, bq.y);
          x.z = fmaf((__uint_as_float(v[4 * j + 2]) - mean) * rstd, gq.z, bq.z);
          x.w = fmaf((__uint_as_float(v[4 * j + 3]) - mean) * rstd, gq.w, bq.w);
          if (j & 1) { s1b += (x.x + x.y) + (x.z + x.w); s2b += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w); }
          else { s1a += (x.x + x.y) + (x.z + x.w); s2a += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w); }
          v[4 * j] = __float_as_uint(x.x * o.park_scale); v[4 * j + 1] = __float_as_uint(x.y * o.park_scale);
          v[4 * j + 2] = __float_as_uint(x.z * o.park_scale); v[4 * j + 3] = __float_as_uint(x.w * o.park_scale);
          if (o.store_x) *reinterpret_cast<float4*>(buf + sw_off(r, j)) = x;
        }
        s1 += s1a + s1b;
        s2 += s2a + s2b;
        tmem_st32(taddr + (ch0 + c) * 32, v);
        if (o.store_x) {
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (elected) { resid_tma_store(tmX, buf, n0 + (ch0 + c) * 32, m0, z); bulk_commit(); }
        }
      }
      tmem_st_wait();
      if (NG == 2) {
        xch[256 + grp * 128 + r] = make_float2(s1, s2);
        named_bar_sync(xbar, 256);
        const float2 o = xch[256 + (grp ^ 1) * 128 + r];
        s1 += o.x; s2 += o.y;
      }
      mean = s1 * inv_n;
      rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + p.eps);
    }
    // ---- final pass: y = LN(X) as bf16 (64-column sub-tiles), optional row mask
    const float* g = sparam + (ln == 2 ? 3 * BN : BN);
    const float* be = sparam + (ln == 2 ? 4 * BN : 2 * BN);
    RTR(10);
    const bool ykeep = (p.y_row_valid == nullptr) || !row_ok || (p.y_row_valid[row] != 0);
    const float unpark = 1.f / o.park_scale;     // exact: park_scale is a power of two (1 / alpha)
    if (!o.store_x) nbuf = 0;                    // nothing of pass 2 went through the ring
#pragma unroll 1
    for (int sub = grp * (BN / 64 / NG); sub < (grp + 1) * (BN / 64 / NG); ++sub, ++nbuf) {
      uint8_t* buf = o.y_smem ? o.y_smem + sub * kBufBytes : ring + (nbuf % R) * kBufBytes;
      if (!o.y_smem && nbuf >= R) { if (elected) bulk_wait_read<R - 1>(); named_bar_sync(bar_id, 128); }
      uint32_t v[64];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        tmem_ld32(taddr + sub * 64, v0);
        tmem_ld32(taddr + sub * 64 + 32, v1);
      }
      tmem_ld_wait();
      const float4* g4 = reinterpret_cast<const float4*>(g + sub * 64);
      const float4* be4 = reinterpret_cast<const float4*>(be + sub * 64);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 ga = g4[2 * j], gb = g4[2 * j + 1], ba = be4[2 * j], bb = be4[2 * j + 1];
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float bbv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float y = fmaf((__uint_as_float(v[8 * j + e]) * unpark - mean) * rstd, gg[e], bbv[e]);
          f[e] = ykeep ? y : 0.f;
        }
        *reinterpret_cast<uint4*>(buf + sw_off(r, j)) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
      fence_proxy_async_smem();
      if (!o.y_smem) {
        named_bar_sync(bar_id, 128);
        if (elected) { resid_tma_store(tmY, buf, sub * 64, m0, z); bulk_commit(); }
      }
    }
  }
  RTR(11);
  tc_fence_before();
  // the staging ring must be drained before anything overwrites it (next tile's residual prefetch); with two groups
  // nobody may restage sparam / xch for the next tile while the other group still reads them
  if (elected) bulk_wait_read<0>();
  if (NG == 1) named_bar_sync(bar_id, 128); else named_bar_sync(xbar, 128 * NG);
  RTR(12);
#undef RTR
}

}  // namespace tc
}  // namespace cfm
