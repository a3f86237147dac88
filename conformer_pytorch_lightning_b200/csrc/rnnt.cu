// RNN-T joint + loss on device (scope row f4): the memory-bound pieces around the joint network's GEMMs
// (joint.py:20-38: out = ffn_out(tanh(enc_ffn(enc) + pred_ffn(pred)))) and the transducer loss that the reference gets
// from torchaudio.functional.rnnt_loss (model.py:95-113; fused log-softmax, blank given, reduction outside).
//   joint_add_tanh        z[b,t,u,:] = tanh(e[b,t,:] + p[b,u,:])                       (joint.py:33-35)
//   joint_tanh_bwd        dpre = dz * (1 - z^2) in place, and the two broadcast reductions
//                         de[b,t,:] = sum_u dpre, dp[b,u,:] = sum_t dpre
//   rnnt_lse_gather       per lattice node (b,t,u): lse over the vocabulary, log p(blank), log p(label_{u+1})
//   rnnt_alpha_beta       forward / backward variables over the (t, u) lattice, anti-diagonal wavefront, one block per
//                         (utterance, direction); nll_b = -(alpha(T-1,U) + log p_blank(T-1,U))
//   rnnt_grad             gradient w.r.t. the LOGITS (log-softmax folded in), written over the logits:
//                         g[v] = scale * ( softmax[v] * exp(alpha+beta-ll) - [v==blank] occ_blank - [v==label] occ_label )
#include "cfm_common.cuh"
#include <math_constants.h>

namespace cfm {
namespace {

__device__ __forceinline__ float log_add2(float a, float b) {
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

template <typename T>
__global__ void __launch_bounds__(256)
joint_add_tanh_kernel(const T* __restrict__ e, const T* __restrict__ p, T* __restrict__ z, int B, int Tn, int U1, int J) {
  // one thread = 8 consecutive channels of one (b,t,u)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int jv = J / 8;
  if (i >= (long long)B * Tn * U1 * jv) return;
  const int c = (int)(i % jv) * 8;
  const long long node = i / jv;
  const int u = (int)(node % U1);
  const long long bt = node / U1;
  const int b = (int)(bt / Tn);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = to_f32(e[bt * J + c + k]) + to_f32(p[((long long)b * U1 + u) * J + c + k]);
    z[node * J + c + k] = from_f32<T>(tanhf(v));
  }
}

// dz (rows, J) -> dpre = dz * (1 - z^2) written over dz
template <typename T>
__global__ void __launch_bounds__(256)
joint_tanh_bwd_kernel(T* dz, const T* __restrict__ z, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float zz = to_f32(z[i]);
  dz[i] = from_f32<T>(to_f32(dz[i]) * (1.f - zz * zz));
}
// out[b, keep, :] = sum over the reduced lattice axis of dpre (B, T, U1, J); reduce_t = 1: sum over t (-> (B,U1,J))
template <typename T>
__global__ void __launch_bounds__(256)
joint_reduce_kernel(const T* __restrict__ dpre, T* __restrict__ out, int B, int Tn, int U1, int J, int reduce_t) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int keep = reduce_t ? U1 : Tn;
  if (i >= (long long)B * keep * J) return;
  const int c = (int)(i % J);
  const long long bk = i / J;
  const int k = (int)(bk % keep), b = (int)(bk / keep);
  float s = 0.f;
  if (reduce_t) {
    for (int t = 0; t < Tn; ++t) s += to_f32(dpre[(((long long)b * Tn + t) * U1 + k) * J + c]);
  } else {
    for (int u = 0; u < U1; ++u) s += to_f32(dpre[(((long long)b * Tn + k) * U1 + u) * J + c]);
  }
  out[i] = from_f32<T>(s);
}

template <typename T>
__global__ void __launch_bounds__(256)
rnnt_lse_gather_kernel(const T* __restrict__ logits, long long ld, int V, int blank, const int* __restrict__ targets, int Umax,
                       int B, int Tn, int U1, float* __restrict__ lse, float* __restrict__ lpb, float* __restrict__ lpl) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * Tn * U1) return;
  const int u = (int)(row % U1);
  const int b = (int)(row / ((long long)Tn * U1));
  const T* x = logits + row * ld;
  float mx = -CUDART_INF_F;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, to_f32(x[v]));
  mx = warp_max(mx);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(to_f32(x[v]) - mx);
  s = warp_sum(s);
  if (lane == 0) {
    const float l = mx + logf(s);
    lse[row] = l;
    lpb[row] = to_f32(x[blank]) - l;
    lpl[row] = (u < Umax) ? to_f32(x[targets[(long long)b * Umax + u]]) - l : -CUDART_INF_F;
  }
}

// grid (B, 2): y = 0 alpha, y = 1 beta.  Anti-diagonal wavefront; alpha / beta (B, Tn, U1) in global memory.
__global__ void __launch_bounds__(256)
rnnt_alpha_beta_kernel(const float* __restrict__ lpb, const float* __restrict__ lpl, const int* __restrict__ t_len,
                       const int* __restrict__ u_len, int Tn, int U1, float* __restrict__ alpha, float* __restrict__ beta,
                       float* __restrict__ nll) {
  const int b = blockIdx.x;
  const bool fwd = blockIdx.y == 0;
  const int Tb = min(t_len[b], Tn), Ub = min(u_len[b], U1 - 1);       // lattice (0..Tb-1) x (0..Ub)
  const long long base = (long long)b * Tn * U1;
  const float* pb = lpb + base;
  const float* pl = lpl + base;
  float* out = (fwd ? alpha : beta) + base;
  if (Tb <= 0) {
    if (fwd && threadIdx.x == 0) nll[b] = CUDART_INF_F;
    return;
  }
  const int ndiag = Tb + Ub;                                          // diagonals d = t + u in [0, Tb - 1 + Ub]
  for (int step = 0; step < ndiag; ++step) {
    const int d = fwd ? step : ndiag - 1 - step;
    for (int u = threadIdx.x; u <= Ub; u += blockDim.x) {
      const int t = d - u;
      if (t < 0 || t >= Tb) continue;
      float v;
      if (fwd) {
        if (t == 0 && u == 0) v = 0.f;
        else {
          const float a = t > 0 ? out[(long long)(t - 1) * U1 + u] + pb[(long long)(t - 1) * U1 + u] : -CUDART_INF_F;
          const float c = u > 0 ? out[(long long)t * U1 + u - 1] + pl[(long long)t * U1 + u - 1] : -CUDART_INF_F;
          v = log_add2(a, c);
        }
      } else {
        if (t == Tb - 1 && u == Ub) v = pb[(long long)t * U1 + u];
        else {
          const float a = t < Tb - 1 ? out[(long long)(t + 1) * U1 + u] + pb[(long long)t * U1 + u] : -CUDART_INF_F;
          const float c = u < Ub ? out[(long long)t * U1 + u + 1] + pl[(long long)t * U1 + u] : -CUDART_INF_F;
          v = log_add2(a, c);
        }
      }
      out[(long long)t * U1 + u] = v;
    }
    __syncthreads();                                                  // (global writes of this block are visible after it)
  }
  if (fwd && threadIdx.x == 0)
    nll[b] = -(out[(long long)(Tb - 1) * U1 + Ub] + pb[(long long)(Tb - 1) * U1 + Ub]);
}

// block per lattice node; dlogits may alias logits
template <typename T>
__global__ void __launch_bounds__(256)
rnnt_grad_kernel(const T* logits, long long ld, int V, int Vp, int blank, const int* __restrict__ targets, int Umax,
                 const int* __restrict__ t_len, const int* __restrict__ u_len, int Tn, int U1, const float* __restrict__ lse,
                 const float* __restrict__ lpb, const float* __restrict__ lpl, const float* __restrict__ alpha,
                 const float* __restrict__ beta, const float* __restrict__ nll, float scale, T* dlogits) {
  const long long row = blockIdx.x;
  const int u = (int)(row % U1);
  const long long bt = row / U1;
  const int t = (int)(bt % Tn), b = (int)(bt / Tn);
  T* out = dlogits + row * ld;
  const int Tb = min(t_len[b], Tn), Ub = min(u_len[b], U1 - 1);
  if (t >= Tb || u > Ub) {
    for (int v = threadIdx.x; v < Vp; v += blockDim.x) out[v] = from_f32<T>(0.f);
    return;
  }
  const float ll = -nll[b];
  const float a = alpha[row];
  const float node = expf(a + beta[row] - ll);                        // occupancy of the node
  float occ_b, occ_l = 0.f;
  if (t == Tb - 1) occ_b = (u == Ub) ? expf(a + lpb[row] - ll) : 0.f;
  else occ_b = expf(a + lpb[row] + beta[row + U1] - ll);
  int label = -1;
  if (u < Ub) { label = targets[(long long)b * Umax + u]; occ_l = expf(a + lpl[row] + beta[row + 1] - ll); }
  const T* x = logits + row * ld;
  const float l = lse[row];
  for (int v = threadIdx.x; v < Vp; v += blockDim.x) {
    float g = 0.f;
    if (v < V) {
      g = node * expf(to_f32(x[v]) - l);
      if (v == blank) g -= occ_b;
      if (v == label) g -= occ_l;
      g *= scale;
    }
    out[v] = from_f32<T>(g);
  }
}

#define CFM_BY_DTYPE(dtype, CALL)                       \
  do {                                                  \
    if ((dtype) == CFM_F32) { using T = float; CALL; }  \
    else { using T = __nv_bfloat16; CALL; }             \
  } while (0)

}  // namespace
}  // namespace cfm

using namespace cfm;

extern "C" int cfm_joint_add_tanh(const void* e, const void* p, void* z, int B, int Tn, int U1, int J, int dtype, void* stream) {
  CFM_CHECK_ARG(e && p && z, "cfm_joint_add_tanh: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_joint_add_tanh: bad dtype");
  CFM_CHECK_ARG(J % 8 == 0, "cfm_joint_add_tanh: join dim %d must be a multiple of 8", J);
  const long long n = (long long)B * Tn * U1 * (J / 8);
  if (n <= 0) return 0;
  CFM_BY_DTYPE(dtype, (joint_add_tanh_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                          (const T*)e, (const T*)p, (T*)z, B, Tn, U1, J)));
  CFM_LAUNCHED_K("joint_add_tanh");
  return 0;
}

extern "C" int cfm_joint_tanh_bwd(void* dz, const void* z, void* de, void* dp, int B, int Tn, int U1, int J, int dtype,
                                  void* stream) {
  CFM_CHECK_ARG(dz && z && de && dp, "cfm_joint_tanh_bwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_joint_tanh_bwd: bad dtype");
  const long long n = (long long)B * Tn * U1 * J;
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CFM_BY_DTYPE(dtype, (joint_tanh_bwd_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((T*)dz, (const T*)z, n)));
  CFM_LAUNCHED_K("joint_tanh_bwd");
  const long long ne = (long long)B * Tn * J, np_ = (long long)B * U1 * J;
  CFM_BY_DTYPE(dtype, (joint_reduce_kernel<T><<<(unsigned)((ne + 255) / 256), 256, 0, st>>>((const T*)dz, (T*)de, B, Tn, U1, J, 0)));
  CFM_BY_DTYPE(dtype, (joint_reduce_kernel<T><<<(unsigned)((np_ + 255) / 256), 256, 0, st>>>((const T*)dz, (T*)dp, B, Tn, U1, J, 1)));
  CFM_LAUNCHED_K("joint_reduce");
  return 0;
}

extern "C" int64_t cfm_rnnt_loss_ws_bytes(int B, int Tn, int U1) {
  return (long long)sizeof(float) * 5LL * B * Tn * U1 + 256;     // lse, lpb, lpl, alpha, beta
}

extern "C" int cfm_rnnt_loss_fwd(const void* logits, int64_t ld, int B, int Tn, int U1, int V, int blank, const int* targets,
                                 int Umax, const int* t_len, const int* u_len, float* nll, void* ws, int dtype, void* stream) {
  CFM_CHECK_ARG(logits && targets && t_len && u_len && nll && ws, "cfm_rnnt_loss_fwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_rnnt_loss_fwd: bad dtype");
  CFM_CHECK_ARG(B >= 0 && Tn > 0 && U1 > 0 && V > 0 && blank >= 0 && blank < V && ld >= V && Umax >= U1 - 1,
                "cfm_rnnt_loss_fwd: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nodes = (long long)B * Tn * U1;
  float* lse = (float*)ws;
  float* lpb = lse + nodes;
  float* lpl = lpb + nodes;
  float* alpha = lpl + nodes;
  float* beta = alpha + nodes;
  CFM_BY_DTYPE(dtype, (rnnt_lse_gather_kernel<T><<<(unsigned)((nodes + 7) / 8), 256, 0, st>>>((const T*)logits, ld, V, blank, targets,
                                                                                             Umax, B, Tn, U1, lse, lpb, lpl)));
  CFM_LAUNCHED_K("rnnt_lse_gather");
  rnnt_alpha_beta_kernel<<<dim3(B, 2), 256, 0, st>>>(lpb, lpl, t_len, u_len, Tn, U1, alpha, beta, nll);
  CFM_LAUNCHED_K("rnnt_alpha_beta");
  return 0;
}

extern "C" int cfm_rnnt_loss_bwd(const void* logits, int64_t ld, int B, int Tn, int U1, int V, int Vp, int blank,
                                 const int* targets, int Umax, const int* t_len, const int* u_len, const float* nll,
                                 const void* ws, float scale, void* dlogits, int dtype, void* stream) {
  CFM_CHECK_ARG(logits && targets && t_len && u_len && nll && ws && dlogits, "cfm_rnnt_loss_bwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_rnnt_loss_bwd: bad dtype");
  CFM_CHECK_ARG(Vp >= V && ld >= Vp, "cfm_rnnt_loss_bwd: bad padded vocabulary");
  if (B == 0) return 0;
  const long long nodes = (long long)B * Tn * U1;
  CFM_CHECK_ARG(nodes < 2147483647LL, "cfm_rnnt_loss_bwd: lattice too large");
  const float* lse = (const float*)ws;
  const float* lpb = lse + nodes;
  const float* lpl = lpb + nodes;
  const float* alpha = lpl + nodes;
  const float* beta = alpha + nodes;
  CFM_BY_DTYPE(dtype, (rnnt_grad_kernel<T><<<(unsigned)nodes, 256, 0, (cudaStream_t)stream>>>(
                          (const T*)logits, ld, V, Vp, blank, targets, Umax, t_len, u_len, Tn, U1, lse, lpb, lpl, alpha, beta, nll,
                          scale, (T*)dlogits)));
  CFM_LAUNCHED_K("rnnt_grad");
  return 0;
}
