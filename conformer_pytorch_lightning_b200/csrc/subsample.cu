// Conv2d sub-sampling front-end (row f1 of the scope table; reference src/convolution.py:52-76):
//     Conv2d(1 -> C, 3, stride 2) + ReLU  ->  Conv2d(C -> C, 3, stride 2) + ReLU  ->  (Linear is a plain cfm_gemm)
// on the bf16 path.  PyTorch/cuDNN needs ~3 ms for this at B = 64 x 10 s (the (B,C,T/2,39) activation alone is
// 1.3 GB in fp32), more than the whole 12-layer stack on the tcgen05 kernels.
//
//   subsample_conv1_kernel   CUDA cores (9 MACs per output).  One warp per output position, lane = 8 channels,
//                            512-byte coalesced bf16 stores.  Output layout is channels-last and PARITY-SPLIT:
//                                P[b][pt][pf][th][fh][c]   with  t1 = 2*th + pt,  f1 = 2*fh + pf
//                            so that the stride-2 gather of the second convolution becomes, per filter tap (i,j), a
//                            dense box of plane (i%2, j%2) shifted by (i/2, j/2) -- i.e. a plain 5-D TMA load.
//   subsample_conv2_kernel   tcgen05 implicit GEMM:  D[(t2,f2), co] = sum_{tap, ci} P[tap-shifted (t2,f2), ci] * W[co, tap, ci]
//                            M tile = TL time rows x F2 frequency bins (6 x 19 = 114 of the 128 MMA rows), N = 256,
//                            K = 9 taps x C in 64-wide blocks.  A tiles come straight from P through TMA (no im2col
//                            buffer), B tiles from the (co, tap*C + ci) K-major weight; persistent, warp-specialised,
//                            TMEM double buffered, bias + ReLU epilogue, bf16 TMA store into O[b][t2][f2][c]
//                            (= the (B*T2, F2*C) row-major A operand of the Linear, whose weight columns the host
//                            permutes from the reference's c*F2+f order to f*C+c).
#include "cfm_common.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace cfm {
namespace {

using namespace tc;

// ------------------------------------------------------------------ conv1 + ReLU
// packed fp32 FMA (Blackwell): two independent fp32 FMAs per instruction
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

// One warp per (b, t1) output row: the three input rows it needs (3 x idim floats) are staged in shared memory once,
// then the warp walks the F1 frequency bins; lane = 8 output channels (16-byte store, 512 B per warp and position).
// 9 broadcast LDS + 36 packed FMAs per position -- the kernel is bound by the 650 MB it writes, not by issue.
constexpr int kC1Warps = 8;
__global__ void __launch_bounds__(kC1Warps * 32)
subsample_conv1_kernel(const float* __restrict__ x,      // (B, Tin, idim) fp32
                       const float* __restrict__ w,      // (C, 9) fp32  [co][i*3+j]
                       const float* __restrict__ bias,   // (C)
                       __nv_bfloat16* __restrict__ P,    // (B, 2, 2, T1h, F1h, C) bf16
                       int B, int Tin, int idim, int C, int T1, int F1, int T1h, int F1h) {
  extern __shared__ float srow[];             // [kC1Warps][3][idim]
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 256 + lane * 8;
  float2 wr[4][9], br[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    br[u] = make_float2(bias[c0 + 2 * u], bias[c0 + 2 * u + 1]);
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[u][k] = make_float2(w[(c0 + 2 * u) * 9 + k], w[(c0 + 2 * u + 1) * 9 + k]);
  }
  float* my = srow + warp * 3 * idim;
  const int n_rows = B * T1;
  for (int row = blockIdx.x * kC1Warps + warp; row < n_rows; row += gridDim.x * kC1Warps) {
    const int b = row / T1, t1 = row % T1;
    const float* xin = x + ((size_t)b * Tin + 2 * t1) * idim;
    __syncwarp();
    for (int i = lane; i < 3 * idim; i += 32) my[i] = __ldg(xin + i);
    __syncwarp();
    const int pt = t1 & 1, th = t1 >> 1;
    __nv_bfloat16* prow0 = P + ((((size_t)b * 2 + pt) * 2 + 0) * T1h + th) * (size_t)F1h * C + c0;
    __nv_bfloat16* prow1 = P + ((((size_t)b * 2 + pt) * 2 + 1) * T1h + th) * (size_t)F1h * C + c0;
    for (int f1 = 0; f1 < F1; ++f1) {
      float in[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) in[i * 3 + j] = my[i * idim + 2 * f1 + j];
      float2 o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        o[u] = br[u];
#pragma unroll
        for (int k = 0; k < 9; ++k) ffma2(o[u], make_float2(in[k], in[k]), wr[u][k]);
      }
      __nv_bfloat16* dst = ((f1 & 1) ? prow1 : prow0) + (size_t)(f1 >> 1) * C;
      *reinterpret_cast<uint4*>(dst) =
          make_uint4(pack_bf16x2(fmaxf(o[0].x, 0.f), fmaxf(o[0].y, 0.f)), pack_bf16x2(fmaxf(o[1].x, 0.f), fmaxf(o[1].y, 0.f)),
                     pack_bf16x2(fmaxf(o[2].x, 0.f), fmaxf(o[2].y, 0.f)), pack_bf16x2(fmaxf(o[3].x, 0.f), fmaxf(o[3].y, 0.f)));
    }
  }
}

// ------------------------------------------------------------------ conv2 + ReLU (implicit GEMM on tcgen05)
constexpr int BK = 64, BN = 256;
constexpr int kABytes = 128 * BK * 2;          // 16 KB slot (TL*F2 rows used)
constexpr int kBBytes = BN * BK * 2;           // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStages = 3;
constexpr int kBuf = 128 * 128;                // staging tile
constexpr int kThreads = 384;
constexpr int kSmemBytes = kStages * kStageBytes + 4 * kBuf + BN * 4 + 256;
static_assert(kSmemBytes <= 232448, "smem budget");

struct Conv2Params {
  const float* bias;
  int B, C, T2, F2, TL;        // TL = time rows per tile (TL*F2 <= 128)
  int tiles_per_utt, n_blocks; // n_blocks = C / 256
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
subsample_conv2_kernel(const __grid_constant__ CUtensorMap tmP,   // P as (C, F1h, T1h, 4 planes, B), box (64, F2, TL, 1, 1)
                       const __grid_constant__ CUtensorMap tmW,   // W (C, 9*C) K-major, box (64, 256)
                       const __grid_constant__ CUtensorMap tmO,   // O as (C, F2, T2, B), box (64, F2, TL, 1)
                       const Conv2Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem + kStages * kStageBytes;
  float* sbias = reinterpret_cast<float*>(ring + 4 * kBuf);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sbias + BN);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.B * p.tiles_per_utt * p.n_blocks;
  const int kb_per_tap = p.C / BK;
  const int kb_count = 9 * kb_per_tap;
  const uint32_t a_bytes = (uint32_t)(p.TL * p.F2) * 128u;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmP); prefetch_tmap(&tmW); prefetch_tmap(&tmO); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 256); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> (n block, utterance, time-tile); n fastest so that neighbouring CTAs share the A tiles in L2
  auto decode = [&](int t, int& nb, int& b, int& t20) {
    nb = t % p.n_blocks;
    const int mt = t / p.n_blocks;
    b = mt / p.tiles_per_utt;
    t20 = (mt % p.tiles_per_utt) * p.TL;
  };

  if (warp == 0) {
    int stage = 0, phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int nb, b, t20;
      decode(t, nb, b, t20);
      for (int kb = 0; kb < kb_count; ++kb) {
        const int tap = kb / kb_per_tap, kc = kb % kb_per_tap;
        const int i = tap / 3, j = tap % 3;
        mbar_wait(empty_bar + stage, phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * kStageBytes;
          mbar_expect_tx(full_bar + stage, a_bytes + kBBytes);
          // plane (i%2, j%2) shifted by (i/2, j/2): rows (t2,f2) of the tile, 64 input channels
          tma_load_5d(sa, &tmP, full_bar + stage, kc * BK, j >> 1, t20 + (i >> 1), (i & 1) * 2 + (j & 1), b);
          tma_load_2d(sa + kABytes, &tmW, full_bar + stage, tap * p.C + kc * BK, nb * BN);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    int stage = 0, phase = 0, it = 0;
    bool have = false;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar + acc, acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < kb_count; ++kb) {
        if (!have) mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        {   // probe the next slot; consumed after this stage's MMAs have been issued
          const int ns = (stage + 1 == kStages) ? 0 : stage + 1;
          have = mbar_test(full_bar + ns, (stage + 1 == kStages) ? (phase ^ 1) : phase);
        }
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(empty_bar + stage);
          if (kb == kb_count - 1) umma_commit(tfull_bar + acc);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // epilogue: two warpgroups x 128 columns, thread = tile row (rows >= TL*F2 are padding and never stored)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int grp = (warp - 4) >> 2;
    const int et = threadIdx.x - 128 - grp * 128;
    const bool elected = (et == 0);
    const int bar_id = 1 + grp;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int sub_cnt = 0, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1, acc_phase = (it >> 1) & 1;
      int nb, b, t20;
      decode(t, nb, b, t20);
      sbias[grp * 128 + et] = p.bias[nb * BN + grp * 128 + et];
      const uint32_t taddr = tmem_base + lane_base + acc * BN;
      mbar_wait(tfull_bar + acc, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ss = 0; ss < 2; ++ss, ++sub_cnt) {
        const int sub = grp * 2 + ss;
        uint8_t* buf = ring + (grp * 2 + (sub_cnt & 1)) * kBuf;
        if (elected) bulk_wait_read<1>();
        named_bar_sync(bar_id, 128);
        uint32_t v[64];
        {
          uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          tmem_ld32(taddr + sub * 64, v0);
          tmem_ld32(taddr + sub * 64 + 32, v1);
        }
        tmem_ld_wait();
        const float* bs = sbias + sub * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[8 * j + e]) + bs[8 * j + e], 0.f);
          *reinterpret_cast<uint4*>(buf + r * 128 + (((j ^ r) & 7) << 4)) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (elected) {
          tma_store_4d(&tmO, buf, nb * BN + sub * 64, 0, t20, b);   // box (64, F2, TL, 1): rows past T2 are clipped
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + acc);
    }
    if (elected) bulk_wait_read<0>();   // the stores only have to be done READING shared memory before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------ conv1 + conv2 in one kernel (C == 256)
// conv1's 653 MB activation is the front-end's whole problem: writing it is bound by the write side of HBM (210 us)
// and conv2 then reads it back.  Here it never exists: for every k-step (filter tap (i,j), 64 input channels) of conv2's
// implicit GEMM, the A tile [128 positions x 64 channels] of relu(conv1) is computed on the spot:
//   converter warps   im2col of the fp32 features for the tap -> bf16 [128 x 16] (one tcgen05 K step; 4 taps share a tile)
//   MMA warp          conv1:  acc1 = im2col . W1_kc^T        (M=128, N=64, K=16, one MMA, two TMEM buffers)
//   converter warps   acc1 -> + b1, ReLU -> bf16 -> the swizzled A slot of the stage ring (where TMA used to put it)
//   MMA warp          conv2:  acc  += A . W2_(tap,kc)^T       (M=128, N=256, K=64), W2 tiles by TMA as before
// conv1 is recomputed for each of the 9 taps (0.3 % more tensor work) and runs two k-steps ahead of conv2, across tile
// boundaries, so the tensor pipe only ever waits for conv2's own operands.  The conv2 accumulator is single-buffered
// (TMEM: 256 + 2 x 64 columns), its bias + ReLU + 4-D TMA store epilogue has its own warpgroup.
constexpr int kFusedThreads = 512;           // 4 control warps + 2 converter warpgroups + 1 epilogue warpgroup
constexpr int kImBytes = 128 * 128;            // im2col tile: 4 K-step slots (taps t, t+1, t+2, t+3 mod 4)
constexpr int kW1Bytes = 64 * 128;             // conv1 filters: row n = channel n of chunk kc, K-step slot kc
constexpr int kBSt = 4, kASt = 2;             // W2 tiles need a deep ring (TMA latency), A tiles are produced locally
constexpr int kFusedSmem = kBSt * kBBytes + kASt * kABytes + kImBytes + kW1Bytes + 2 * kBuf + 2 * BN * 4 + 512;
static_assert(kFusedSmem <= 232448, "smem budget");

struct FusedParams {
  const float* x; const float* w1; const float* b1; const float* b2;
  int B, Tin, idim, T2, F2, TL, tiles_per_utt;
};

// ATMEM: relu(conv1) goes straight back into tensor memory (tcgen05.st of the packed bf16 row, 32 columns per k-step, two
// slots in the 128 columns the accumulators leave free) and conv2 reads its A operand from there (tcgen05.mma with A in
// TMEM): no shared-memory A tile, no st.shared + fence.proxy.async per k-step in the converter warps, and one operand
// less on the shared-memory port.
template <bool ATMEM>
__global__ void __launch_bounds__(kFusedThreads, 1)
subsample_fused_kernel(const __grid_constant__ CUtensorMap tmW,   // W2 (C, 9*C) K-major, box (64, 256)
                       const __grid_constant__ CUtensorMap tmO,   // O as (C, F2, T2, B), box (64, F2, TL, 1)
                       const FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;                                   // kBSt x 32 KB
  uint8_t* sAt = sB + kBSt * kBBytes;                   // kASt x 16 KB
  uint8_t* sIm = sAt + kASt * kABytes;
  uint8_t* sW1 = sIm + kImBytes;
  uint8_t* ring = sW1 + kW1Bytes;                       // 2 output staging tiles
  float* sb1 = reinterpret_cast<float*>(ring + 2 * kBuf);
  float* sb2 = sb1 + BN;
  // ATMEM: the two shared-memory A slots (32 KB, directly behind the W2 ring) are not needed: a fifth W2 stage lives there
  constexpr int NB = ATMEM ? kBSt + 1 : kBSt;
  static_assert(kASt * kABytes == kBBytes, "the fifth W2 stage takes exactly the A slots");
  uint64_t* b_full = reinterpret_cast<uint64_t*>(sb2 + BN);   // [NB] W2 tile landed
  uint64_t* b_empty = b_full + (kBSt + 1);              // [NB] conv2 MMAs that read the slot retired
  uint64_t* a_full = b_empty + (kBSt + 1);              // [kASt] A tile written by converter group g & 1 (128 arrivals)
  uint64_t* a_empty = a_full + kASt;                    // [kASt]
  uint64_t* im_full = a_empty + kASt;                   // [4] im2col slot written (128 arrivals)
  uint64_t* im_empty = im_full + 4;                     // [4] conv1 MMAs of the tap retired
  uint64_t* acc1_full = im_empty + 4;                   // [2]
  uint64_t* acc1_empty = acc1_full + 2;                 // [2] 128 arrivals
  uint64_t* tfull = acc1_empty + 2;
  uint64_t* tempty = tfull + 1;                         // 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = p.B * p.tiles_per_utt;
  const int n_my = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
  const int n_ks = n_my * 36;                           // k-steps of this CTA, one stream across its tiles
  const int rows_used = p.TL * p.F2;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmW); prefetch_tmap(&tmO); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NB; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int s = 0; s < kASt; ++s) { mbar_init(a_full + s, 128); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(im_full + s, 64); mbar_init(im_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc1_full + s, 1); mbar_init(acc1_empty + s, 128); }
    mbar_init(tfull, 1); mbar_init(tempty, 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  pdl_launch_dependents();
  // conv1 filters (C x 9 fp32) -> bf16 B operand, biases -> smem (parameters: not produced by a preceding kernel)
  for (int co = threadIdx.x; co < BN; co += kFusedThreads) {
    float wv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wv[k] = __ldg(p.w1 + co * 9 + k);
    const int n = co & 63, kc = co >> 6;
    uint8_t* rowp = sW1 + n * 128;
    *reinterpret_cast<uint4*>(rowp + ((((2 * kc) ^ n) & 7) << 4)) =
        make_uint4(pack_bf16x2(wv[0], wv[1]), pack_bf16x2(wv[2], wv[3]), pack_bf16x2(wv[4], wv[5]), pack_bf16x2(wv[6], wv[7]));
    *reinterpret_cast<uint4*>(rowp + ((((2 * kc + 1) ^ n) & 7) << 4)) = make_uint4(pack_bf16x2(wv[8], 0.f), 0u, 0u, 0u);
    sb1[co] = __ldg(p.b1 + co);
    sb2[co] = __ldg(p.b2 + co);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc1 = tmem_base + BN;            // 2 x 64 columns
  const uint32_t tmem_a = tmem_base + BN + 128;         // ATMEM: 2 x 32 columns (A operand of conv2, 64 packed bf16 per row)

  if (warp == 0) {
    // ===================== TMA producer: W2 tiles =====================
    for (int g = 0; g < n_ks; ++g) {
      const int ks = g % 36, tap = ks >> 2, kc = ks & 3;
      const int bs = g % NB;
      mbar_wait(b_empty + bs, ((g / NB) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(b_full + bs, kBBytes);
        tma_load_2d(sB + bs * kBBytes, &tmW, b_full + bs, tap * BN + kc * BK, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc2 = umma_idesc_bf16(128, BN);
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64);
    const uint64_t dim0 = umma_desc_sw128(smem_u32(sIm)), dw1 = umma_desc_sw128(smem_u32(sW1));
    auto conv1 = [&](int g) {                           // k-step g of the CTA's stream
      const int tapg = g >> 2, kc = g & 3;              // global tap counter (slot = tapg & 3)
      const int ab = g & 1;
      if (kc == 0) { mbar_wait(im_full + (tapg & 3), (tapg >> 2) & 1); }
      mbar_wait(acc1_empty + ab, ((g >> 1) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        umma_bf16(tmem_acc1 + ab * 64, dim0 + 2 * (tapg & 3), dw1 + 2 * kc, idesc1, 0);
        umma_commit(acc1_full + ab);
        if (kc == 3) umma_commit(im_empty + (tapg & 3));
      }
      __syncwarp();
    };
    if (n_ks > 0) { conv1(0); conv1(1); }
    for (int g = 0; g < n_ks; ++g) {
      const int ks = g % 36, tl = g / 36;
      const int as = g & 1, bs = g % NB;
      // conv1 of k-step g+2 first: its TMEM buffer was released as soon as the converter had LOADED conv1(g), so it runs while
      // the converter is still turning conv1(g) into the A operand of conv2(g) (issued after conv2(g) it used to sit behind
      // that whole conversion: the converter chain, not the tensor pipe, set the pace)
      if (g + 2 < n_ks) conv1(g + 2);
      if (ks == 0) { mbar_wait(tempty, (tl & 1) ^ 1); }  // the epilogue has drained the previous tile's accumulator
      mbar_wait(a_full + as, (g >> 1) & 1);
      mbar_wait(b_full + bs, (g / NB) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_sw128(smem_u32(sAt + as * kABytes)), db = umma_desc_sw128(smem_u32(sB + bs * kBBytes));
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if constexpr (ATMEM) umma_bf16_ts(tmem_base, tmem_a + as * 32 + k * 8, db + 2 * k, idesc2, (ks | k) != 0);
          else umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc2, (ks | k) != 0);
        }
        umma_commit(a_empty + as);
        umma_commit(b_empty + bs);
        if (ks == 35) umma_commit(tfull);
      }
      __syncwarp();
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== im2col producers (64 threads, two tile rows each): the 3 x 3 fp32 patch of every position for
    //                       tap `tapg` -> bf16 K-step slot tapg & 3 of the [128 x 16 x 4] im2col tile.  They run up to four taps
    //                       ahead of conv1 (im_empty), so the global-load latency of the patches is off the converters'
    //                       critical path (it used to cost them ~1.5 k cycles per tap) =====================
    const int n_taps = n_my * 9;
    const int tid = threadIdx.x - 64;
    for (int tapg = 0; tapg < n_taps; ++tapg) {
      const int tl = tapg / 9, tap = tapg - tl * 9;
      const int t = (int)blockIdx.x + tl * (int)gridDim.x;
      const int b = t / p.tiles_per_utt, t20 = (t % p.tiles_per_utt) * p.TL;
      const int i = tap / 3, j = tap - i * 3;
      const int slot = tapg & 3;
      float v[2][9];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = tid + 64 * h;
        const int dt2 = r / p.F2, f2 = r - dt2 * p.F2;
        const int t2 = t20 + dt2;
        if (r < rows_used && t2 < p.T2) {
          const float* xin = p.x + ((size_t)b * p.Tin + 4 * t2 + 2 * i) * p.idim + 4 * f2 + 2 * j;
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) v[h][a * 3 + c] = __ldg(xin + a * p.idim + c);
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) v[h][k] = 0.f;
        }
      }
      mbar_wait(im_empty + slot, ((tapg >> 2) & 1) ^ 1);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = tid + 64 * h;
        uint8_t* rowp = sIm + r * 128;
        *reinterpret_cast<uint4*>(rowp + ((((2 * slot) ^ r) & 7) << 4)) =
            make_uint4(pack_bf16x2(v[h][0], v[h][1]), pack_bf16x2(v[h][2], v[h][3]), pack_bf16x2(v[h][4], v[h][5]), pack_bf16x2(v[h][6], v[h][7]));
        *reinterpret_cast<uint4*>(rowp + ((((2 * slot + 1) ^ r) & 7) << 4)) = make_uint4(pack_bf16x2(v[h][8], 0.f), 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
      mbar_arrive(im_full + slot);
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== converter warps: thread = tile row (position); warpgroup cg takes the k-steps g = cg mod 2
    //                       (= TMEM buffer acc1[cg]) and builds the im2col tiles of the taps tg = cg mod 2 =====================
    const int cg = (warp - 4) >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    for (int g = cg; g < n_ks; g += 2) {
      const int kc = g & 3, ab = g & 1;
      mbar_wait(acc1_full + ab, (g >> 1) & 1);
      tc_fence_after();
      uint32_t v[64];
      {
        uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        tmem_ld32(tmem_acc1 + lane_base + ab * 64, v0);
        tmem_ld32(tmem_acc1 + lane_base + ab * 64 + 32, v1);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(acc1_empty + ab);
      const float* bs = sb1 + kc * 64;
      uint4 pk[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[8 * jj + e]) + bs[8 * jj + e], 0.f);
        pk[jj] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
      mbar_wait(a_empty + cg, ((g >> 1) & 1) ^ 1);       // conv2 has finished with the A tile that lived here (slot g & 1 = cg)
      if constexpr (ATMEM) {
        tc_fence_after();
        uint32_t pr[32];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) { pr[4 * jj] = pk[jj].x; pr[4 * jj + 1] = pk[jj].y; pr[4 * jj + 2] = pk[jj].z; pr[4 * jj + 3] = pk[jj].w; }
        tmem_st32(tmem_a + lane_base + cg * 32, pr);
        tmem_st_wait();
        tc_fence_before();
      } else {
        uint8_t* a = sAt + cg * kABytes;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<uint4*>(a + r * 128 + (((jj ^ r) & 7) << 4)) = pk[jj];
        fence_proxy_async_smem();
      }
      mbar_arrive(a_full + cg);
    }
  } else if (warp >= 12) {
    // ===================== output epilogue: bias + ReLU + 4-D TMA store, thread = tile row =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 384;
    const bool elected = (et == 0);
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int sub_cnt = 0;
    for (int tl = 0; tl < n_my; ++tl) {
      const int t = (int)blockIdx.x + tl * (int)gridDim.x;
      const int b = t / p.tiles_per_utt, t20 = (t % p.tiles_per_utt) * p.TL;
      mbar_wait(tfull, tl & 1);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub, ++sub_cnt) {
        uint8_t* buf = ring + (sub_cnt & 1) * kBuf;
        if (elected) bulk_wait_read<1>();
        named_bar_sync(1, 128);
        uint32_t v[64];
        {
          uint32_t (&v0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&v1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          tmem_ld32(tmem_base + lane_base + sub * 64, v0);
          tmem_ld32(tmem_base + lane_base + sub * 64 + 32, v1);
        }
        tmem_ld_wait();
        const float* bs = sb2 + sub * 64;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[8 * jj + e]) + bs[8 * jj + e], 0.f);
          *reinterpret_cast<uint4*>(buf + r * 128 + (((jj ^ r) & 7) << 4)) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (elected) {
          tma_store_4d(&tmO, buf, sub * 64, 0, t20, b);   // box (64, F2, TL, 1): rows past T2 are clipped
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
    if (elected) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int make_map_nd(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box) {
  EncodeTiledFn fn = tc::encode_tiled_fn();
  CFM_CHECK_ARG(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CFM_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %d) failed with CUresult %d", rank, (int)r);
  return 0;
}

}  // namespace
}  // namespace cfm

extern "C" int64_t cfm_subsample_ws_bytes(int B, int Tin, int idim, int C) {
  const int T1 = (Tin - 3) / 2 + 1, F1 = (idim - 3) / 2 + 1;
  const int64_t T1h = (T1 + 1) / 2, F1h = (F1 + 1) / 2;
  return (int64_t)B * 4 * T1h * F1h * C * 2;
}

extern "C" int cfm_subsample_conv(const float* x, int B, int Tin, int idim, const float* w1, const float* b1,
                                  const void* w2, const float* b2, int C, void* ws, void* out, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(x && w1 && b1 && w2 && b2 && ws && out, "cfm_subsample_conv: null pointer");
  CFM_CHECK_ARG(C % 256 == 0, "cfm_subsample_conv: C=%d must be a multiple of 256", C);
  CFM_CHECK_ARG(Tin >= 7 && idim >= 7, "cfm_subsample_conv: input too small");
  const int T1 = (Tin - 3) / 2 + 1, F1 = (idim - 3) / 2 + 1;
  const int T2 = (T1 - 3) / 2 + 1, F2 = (F1 - 3) / 2 + 1;
  CFM_CHECK_ARG(F2 >= 1 && F2 <= 128, "cfm_subsample_conv: F2=%d unsupported", F2);
  if (B <= 0) return 0;
  const int T1h = (T1 + 1) / 2, F1h = (F1 + 1) / 2;
  const int TL = 128 / F2;
  cudaStream_t st = (cudaStream_t)stream;
  // C == 256: conv1 is computed on the fly inside conv2 (CFM_B200_FRONTEND=unfused keeps the two-kernel path)
  static const bool fe_unfused = env_is("CFM_B200_FRONTEND", "unfused");
  if (!fe_unfused && C == 256 && tc::encode_tiled_fn() != nullptr) {
    CUtensorMap tmW, tmO;
    int rc;
    {
      const uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)C};
      const uint64_t str[1] = {(uint64_t)9 * C * 2};
      const uint32_t box[2] = {64, 256};
      if ((rc = make_map_nd(&tmW, w2, 2, dims, str, box)) != 0) return rc;
    }
    {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)F2, (uint64_t)T2, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)F2 * C * 2, (uint64_t)T2 * F2 * C * 2};
      const uint32_t box[4] = {64, (uint32_t)F2, (uint32_t)TL, 1};
      if ((rc = make_map_nd(&tmO, out, 4, dims, str, box)) != 0) return rc;
    }
    FusedParams fp{x, w1, b1, b2, B, Tin, idim, T2, F2, TL, (T2 + TL - 1) / TL};
    static const bool a_smem = env_is("CFM_B200_SUBSAMPLE_ATMEM", "0");
    CFM_SMEM_OPT_IN(subsample_fused_kernel<true>, kFusedSmem);
    CFM_SMEM_OPT_IN(subsample_fused_kernel<false>, kFusedSmem);
    const int total = B * fp.tiles_per_utt;
    const int grid = total < num_sms() ? total : num_sms();
    if (a_smem) CFM_CUDA_OK(launch_pdl(subsample_fused_kernel<false>, dim3(grid), dim3(kFusedThreads), (size_t)kFusedSmem, st, 1, tmW, tmO, fp));
    else CFM_CUDA_OK(launch_pdl(subsample_fused_kernel<true>, dim3(grid), dim3(kFusedThreads), (size_t)kFusedSmem, st, 1, tmW, tmO, fp));
    CFM_LAUNCHED_K("subsample_fused");
    return 0;
  }
  {
    const int n_rows = B * T1;
    const int blocks = std::min((n_rows + kC1Warps - 1) / kC1Warps, num_sms() * 8);
    dim3 grid(blocks, C / 256);
    const size_t sm = (size_t)kC1Warps * 3 * idim * sizeof(float);
    CFM_CHECK_ARG(sm <= 48 * 1024, "cfm_subsample_conv: idim=%d too large", idim);
    CFM_CUDA_OK(launch_pdl(subsample_conv1_kernel, grid, dim3(kC1Warps * 32), sm, st, 1, x, w1, b1, (__nv_bfloat16*)ws, B, Tin,
                           idim, C, T1, F1, T1h, F1h));
    CFM_LAUNCHED_K("subsample_conv1");
  }
  CUtensorMap tmP, tmW, tmO;
  int rc;
  {
    const uint64_t dims[5] = {(uint64_t)C, (uint64_t)F1h, (uint64_t)T1h, 4, (uint64_t)B};
    const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)F1h * C * 2, (uint64_t)T1h * F1h * C * 2, (uint64_t)4 * T1h * F1h * C * 2};
    const uint32_t box[5] = {64, (uint32_t)F2, (uint32_t)TL, 1, 1};
    if ((rc = make_map_nd(&tmP, ws, 5, dims, str, box)) != 0) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)C};
    const uint64_t str[1] = {(uint64_t)9 * C * 2};
    const uint32_t box[2] = {64, 256};
    if ((rc = make_map_nd(&tmW, w2, 2, dims, str, box)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)F2, (uint64_t)T2, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)F2 * C * 2, (uint64_t)T2 * F2 * C * 2};
    const uint32_t box[4] = {64, (uint32_t)F2, (uint32_t)TL, 1};
    if ((rc = make_map_nd(&tmO, out, 4, dims, str, box)) != 0) return rc;
  }
  Conv2Params p{b2, B, C, T2, F2, TL, (T2 + TL - 1) / TL, C / 256};
  CFM_SMEM_OPT_IN(subsample_conv2_kernel, kSmemBytes);
  const int total = B * p.tiles_per_utt * p.n_blocks;
  const int grid = total < num_sms() ? total : num_sms();
  CFM_CUDA_OK(launch_pdl(subsample_conv2_kernel, dim3(grid), dim3(kThreads), kSmemBytes, st, 1, tmP, tmW, tmO, p));
  CFM_LAUNCHED_K("subsample_conv2");
  return 0;
}
