// CUDA-core flash-style attention (online softmax, no T x T materialisation) with the reference's
// mask semantics (attention.py:84-97): masked -> -inf -> softmax -> masked probabilities = 0, a fully
// masked row yields 0.  fp32 accumulation everywhere; engine of the fp32 path and of the streaming
// (B = 1, chunk of 16 queries) path.  Head dim fixed at 64.
#include "cfm_common.cuh"
#include <math_constants.h>

namespace cfm {
namespace {

// query rows per block = 4 warps x RPW rows (RPW = 8; 2 for streaming chunks, whose 16 rows then spread over 2 x H blocks
// of 4 busy warps instead of 1 x H blocks with two idle warps)
constexpr int KT = 64;      // keys per shared-memory tile
constexpr int DK = 64;
constexpr int KSTR = DK + 1;  // padded K row stride (bank-conflict free column reads)

template <typename T, int RPW>
__global__ void __launch_bounds__(128)
attention_simt_kernel(const T* __restrict__ q, int64_t q_bs, int64_t q_ts, const T* __restrict__ k, int64_t k_bs,
                      int64_t k_ts, const T* __restrict__ v, int64_t v_bs, int64_t v_ts, T* __restrict__ out,
                      int H, int Tq, int Tk, const uint8_t* __restrict__ mask, int64_t mask_bs, int64_t mask_rs,
                      const float* __restrict__ key_bias, float scale) {
  constexpr int QR = 4 * RPW;
  __shared__ float sq[QR][DK];
  __shared__ float sk[KT][KSTR];
  __shared__ float sv[KT][DK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = blockIdx.x * QR;
  const int h = blockIdx.y, b = blockIdx.z;

  for (int idx = threadIdx.x; idx < QR * DK; idx += blockDim.x) {
    const int r = idx / DK, c = idx % DK;
    const int i = i0 + r;
    sq[r][c] = (i < Tq) ? to_f32(q[b * q_bs + i * q_ts + h * DK + c]) : 0.f;
  }

  float m[RPW], l[RPW], o0[RPW], o1[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) { m[r] = -CUDART_INF_F; l[r] = 0.f; o0[r] = 0.f; o1[r] = 0.f; }

  for (int j0 = 0; j0 < Tk; j0 += KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < KT * DK; idx += blockDim.x) {
      const int r = idx / DK, c = idx % DK;
      const int j = j0 + r;
      const bool in = j < Tk;
      sk[r][c] = in ? to_f32(k[b * k_bs + j * k_ts + h * DK + c]) : 0.f;
      sv[r][c] = in ? to_f32(v[b * v_bs + j * v_ts + h * DK + c]) : 0.f;
    }
    __syncthreads();
    const int ja = j0 + lane, jb = j0 + lane + 32;
    const float kb_a = (key_bias && ja < Tk) ? key_bias[((size_t)b * H + h) * Tk + ja] : 0.f;
    const float kb_b = (key_bias && jb < Tk) ? key_bias[((size_t)b * H + h) * Tk + jb] : 0.f;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = warp * RPW + r;
      const int i = i0 + row;
      if (i >= Tq) continue;   // warp-uniform
      float sa = 0.f, sb = 0.f;
#pragma unroll 16
      for (int c = 0; c < DK; ++c) {
        const float qv = sq[row][c];
        sa = fmaf(qv, sk[lane][c], sa);
        sb = fmaf(qv, sk[lane + 32][c], sb);
      }
      sa = (sa + kb_a) * scale;
      sb = (sb + kb_b) * scale;
      bool va = ja < Tk, vb = jb < Tk;
      if (mask != nullptr) {
        const uint8_t* mr = mask + b * mask_bs + i * mask_rs;
        va = va && (mr[ja < Tk ? ja : 0] != 0);
        vb = vb && (mr[jb < Tk ? jb : 0] != 0);
      }
      if (!va) sa = -CUDART_INF_F;
      if (!vb) sb = -CUDART_INF_F;
      const float mt = warp_max(fmaxf(sa, sb));
      const float mn = fmaxf(m[r], mt);
      if (mn == -CUDART_INF_F) continue;   // nothing visible so far (warp-uniform)
      const float corr = (m[r] == -CUDART_INF_F) ? 0.f : expf(m[r] - mn);
      const float pa = va ? expf(sa - mn) : 0.f;
      const float pb = vb ? expf(sb - mn) : 0.f;
      l[r] = l[r] * corr + warp_sum(pa + pb);
      float a0 = o0[r] * corr, a1 = o1[r] * corr;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj = __shfl_sync(0xffffffffu, pa, j);
        a0 = fmaf(pj, sv[j][lane], a0);
        a1 = fmaf(pj, sv[j][lane + 32], a1);
      }
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float pj = __shfl_sync(0xffffffffu, pb, j);
        a0 = fmaf(pj, sv[j + 32][lane], a0);
        a1 = fmaf(pj, sv[j + 32][lane + 32], a1);
      }
      o0[r] = a0; o1[r] = a1; m[r] = mn;
    }
  }

#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int i = i0 + warp * RPW + r;
    if (i >= Tq) continue;
    const float inv = (l[r] > 0.f) ? 1.0f / l[r] : 0.f;   // fully masked row -> 0 (attention.py:92)
    T* orow = out + ((size_t)b * Tq + i) * H * DK + h * DK;
    orow[lane] = from_f32<T>(o0[r] * inv);
    orow[lane + 32] = from_f32<T>(o1[r] * inv);
  }
}

}  // namespace

int attention_simt(const void* q, int64_t q_bs, int64_t q_ts, const void* k, int64_t k_bs, int64_t k_ts,
                   const void* v, int64_t v_bs, int64_t v_ts, void* out, int B, int H, int Tq, int Tk,
                   const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, const float* key_bias, float scale,
                   int dtype, cudaStream_t st) {
  CFM_CHECK_ARG(H <= 65535 && B <= 65535, "cfm_attention: B/H too large for the grid");
  const int rpw = (Tq <= 32) ? 2 : 8;
  dim3 grid((Tq + 4 * rpw - 1) / (4 * rpw), H, B);
#define CFM_ATTN_SIMT(TT, R)                                                                                              \
  attention_simt_kernel<TT, R><<<grid, 128, 0, st>>>((const TT*)q, q_bs, q_ts, (const TT*)k, k_bs, k_ts, (const TT*)v, v_bs, \
                                                     v_ts, (TT*)out, H, Tq, Tk, mask, mask_bs, mask_rs, key_bias, scale)
  if (dtype == CFM_F32) {
    if (rpw == 2) CFM_ATTN_SIMT(float, 2); else CFM_ATTN_SIMT(float, 8);
  } else {
    if (rpw == 2) CFM_ATTN_SIMT(__nv_bfloat16, 2); else CFM_ATTN_SIMT(__nv_bfloat16, 8);
  }
#undef CFM_ATTN_SIMT
  CFM_LAUNCHED_K("attention_simt");
  return 0;
}

}  // namespace cfm
