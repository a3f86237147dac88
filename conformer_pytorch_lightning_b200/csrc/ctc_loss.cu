// CTC loss forward / backward on device (scope row f2): log-softmax statistics over the vocabulary, the alpha / beta
// recursions in log space, and the gradient with respect to the LOGITS (log_softmax folded in) -- what
// CTCDecoder.forward computes with `logits.log_softmax(2)` + nn.CTCLoss(reduction='sum') (decoder.py:18-23 of the
// reference), blank = 0.  The (frames x vocab) log-probability matrix is never materialised: per frame only its
// logsumexp is kept, and the recursions run on a gathered (frames x 2L+1) table of the label log-probabilities.
//   ctc_lse      warp per frame: lse[row] = logsumexp_v logits[row][v]
//   ctc_gather   lpe[b][t][s] = logits[b,t][l'_s] - lse[b,t]       l' = blank-interleaved label sequence
//   ctc_alpha_beta  one block per (utterance, direction): S = 2L+1 states in parallel, T sequential steps
//   ctc_grad     block per frame: dlogits[v] = scale * (softmax[v] - occupancy[v]), occupancy accumulated per label in
//                shared memory (repeated labels / the L+1 blanks collide on the same vocabulary entry)
#include "cfm_common.cuh"
#include <math_constants.h>

namespace cfm {
namespace {

__device__ __forceinline__ float log_add(float a, float b) {
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}

template <typename T>
__global__ void __launch_bounds__(256)
ctc_lse_kernel(const T* __restrict__ logits, long long ld, int rows, int V, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const T* x = logits + (size_t)row * ld;
  float mx = -CUDART_INF_F;
  for (int v = lane; v < V; v += 32) mx = fmaxf(mx, to_f32(x[v]));
  mx = warp_max(mx);
  float s = 0.f;
  for (int v = lane; v < V; v += 32) s += expf(to_f32(x[v]) - mx);
  s = warp_sum(s);
  if (lane == 0) lse[row] = mx + logf(s);
}

template <typename T>
__global__ void __launch_bounds__(256)
ctc_gather_kernel(const T* __restrict__ logits, long long ld, const float* __restrict__ lse, const int* __restrict__ labels,
                  int Lmax, const int* __restrict__ lab_len, int B, int Tlen, int Sp, float* __restrict__ lpe) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Tlen * Sp) return;
  const int s = (int)(i % Sp);
  const long long bt = i / Sp;
  const int b = (int)(bt / Tlen);
  const int S = 2 * lab_len[b] + 1;
  float v = -CUDART_INF_F;
  if (s < S) {
    const int lab = (s & 1) ? labels[(size_t)b * Lmax + (s >> 1)] : 0;
    v = to_f32(logits[(size_t)bt * ld + lab]) - lse[bt];
  }
  lpe[i] = v;
}

// grid (B, 2): y = 0 alpha (forward in time), y = 1 beta (backward).  Dynamic smem: 2 * Sp floats.
__global__ void __launch_bounds__(256)
ctc_alpha_beta_kernel(const float* __restrict__ lpe, const int* __restrict__ labels, int Lmax, const int* __restrict__ in_len,
                      const int* __restrict__ lab_len, int Tlen, int Sp, float* __restrict__ alpha, float* __restrict__ beta,
                      float* __restrict__ nll) {
  extern __shared__ float sh[];
  float* prev = sh;
  float* cur = sh + Sp;
  const int b = blockIdx.x;
  const bool fwd = blockIdx.y == 0;
  const int Tb = min(in_len[b], Tlen), L = lab_len[b], S = 2 * L + 1;
  const float* lp = lpe + (size_t)b * Tlen * Sp;
  float* out = (fwd ? alpha : beta) + (size_t)b * Tlen * Sp;
  const int* lab = labels + (size_t)b * Lmax;
  if (Tb <= 0) {
    if (fwd && threadIdx.x == 0) nll[b] = (L == 0) ? 0.f : CUDART_INF_F;
    return;
  }
  for (int step = 0; step < Tb; ++step) {
    const int t = fwd ? step : Tb - 1 - step;
    for (int s = threadIdx.x; s < Sp; s += blockDim.x) {
      float v = -CUDART_INF_F;
      if (s < S) {
        if (step == 0) {
          if (fwd ? (s <= 1) : (s >= S - 2)) v = lp[(size_t)t * Sp + s];
        } else if (fwd) {
          float a = prev[s];
          if (s >= 1) a = log_add(a, prev[s - 1]);
          if (s >= 2 && (s & 1) && lab[s >> 1] != lab[(s >> 1) - 1]) a = log_add(a, prev[s - 2]);
          v = a + lp[(size_t)t * Sp + s];
        } else {
          float a = prev[s];
          if (s + 1 < S) a = log_add(a, prev[s + 1]);
          if (s + 2 < S && (s & 1) && lab[s >> 1] != lab[(s >> 1) + 1]) a = log_add(a, prev[s + 2]);
          v = a + lp[(size_t)t * Sp + s];
        }
      }
      cur[s] = v;
      out[(size_t)t * Sp + s] = v;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (fwd && threadIdx.x == 0) {
    float ll = prev[S - 1];
    if (S >= 2) ll = log_add(ll, prev[S - 2]);
    nll[b] = -ll;
  }
}

// block per frame (b, t); dynamic smem: V floats.  dlogits may alias logits.
template <typename T>
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const T* logits, long long ld, const float* __restrict__ lse, const float* __restrict__ lpe,
                const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ nll,
                const int* __restrict__ labels, int Lmax, const int* __restrict__ in_len, const int* __restrict__ lab_len,
                int Tlen, int Sp, int V, int Vp, float scale, T* dlogits) {
  extern __shared__ float occ[];
  const int bt = blockIdx.x, b = bt / Tlen, t = bt % Tlen;
  T* out = dlogits + (size_t)bt * ld;
  if (t >= in_len[b]) {
    for (int v = threadIdx.x; v < Vp; v += blockDim.x) out[v] = from_f32<T>(0.f);
    return;
  }
  for (int v = threadIdx.x; v < V; v += blockDim.x) occ[v] = 0.f;
  __syncthreads();
  const int S = 2 * lab_len[b] + 1;
  const float nl = nll[b];
  const size_t base = (size_t)bt * Sp;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int lab = (s & 1) ? labels[(size_t)b * Lmax + (s >> 1)] : 0;
    const float lg = alpha[base + s] + beta[base + s] - lpe[base + s] + nl;     // log occupancy of state s
    if (lg > -CUDART_INF_F) atomicAdd(occ + lab, expf(lg));
  }
  __syncthreads();
  const T* x = logits + (size_t)bt * ld;
  const float l = lse[bt];
  for (int v = threadIdx.x; v < Vp; v += blockDim.x) {
    float g = 0.f;
    if (v < V) g = scale * (expf(to_f32(x[v]) - l) - occ[v]);
    out[v] = from_f32<T>(g);
  }
}

}  // namespace
}  // namespace cfm

using namespace cfm;

extern "C" int64_t cfm_ctc_loss_ws_bytes(int B, int T, int Lmax) {
  const long long Sp = ((2LL * Lmax + 1) + 7) / 8 * 8;
  // lse (B*T) + lpe, alpha, beta (B*T*Sp each)
  return (long long)sizeof(float) * ((long long)B * T + 3LL * B * T * Sp) + 256;
}

// forward: nll[b] = -log p(labels_b | logits_b).  logits (B*T, ld) in `dtype`, V valid columns; labels (B, Lmax) int32;
// in_len, lab_len (B) int32; ws as sized by cfm_ctc_loss_ws_bytes (kept for the backward call).
extern "C" int cfm_ctc_loss_fwd(const void* logits, int64_t ld, int B, int T, int V, const int* labels, int Lmax,
                                const int* in_len, const int* lab_len, float* nll, void* ws, int dtype, void* stream) {
  CFM_CHECK_ARG(logits && labels && in_len && lab_len && nll && ws, "cfm_ctc_loss_fwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_ctc_loss_fwd: bad dtype");
  CFM_CHECK_ARG(B >= 0 && T >= 0 && V > 0 && Lmax >= 0 && ld >= V, "cfm_ctc_loss_fwd: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int Sp = ((2 * Lmax + 1) + 7) / 8 * 8;
  float* lse = (float*)ws;
  float* lpe = lse + ((size_t)B * T + 63) / 64 * 64;
  float* alpha = lpe + (size_t)B * T * Sp;
  float* beta = alpha + (size_t)B * T * Sp;
  const int rows = B * T;
  if (rows > 0) {
    if (dtype == CFM_F32) ctc_lse_kernel<float><<<(rows + 7) / 8, 256, 0, st>>>((const float*)logits, ld, rows, V, lse);
    else ctc_lse_kernel<__nv_bfloat16><<<(rows + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)logits, ld, rows, V, lse);
    CFM_LAUNCHED_K("ctc_lse");
    const long long n = (long long)rows * Sp;
    const int blocks = (int)((n + 255) / 256);
    if (dtype == CFM_F32)
      ctc_gather_kernel<float><<<blocks, 256, 0, st>>>((const float*)logits, ld, lse, labels, Lmax, lab_len, B, T, Sp, lpe);
    else
      ctc_gather_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)logits, ld, lse, labels, Lmax, lab_len, B, T, Sp, lpe);
    CFM_LAUNCHED_K("ctc_gather");
  }
  CFM_CHECK_ARG(2 * Sp * sizeof(float) <= 48 * 1024, "cfm_ctc_loss_fwd: label length %d too large", Lmax);
  ctc_alpha_beta_kernel<<<dim3(B, 2), 256, 2 * Sp * sizeof(float), st>>>(lpe, labels, Lmax, in_len, lab_len, T, Sp, alpha, beta, nll);
  CFM_LAUNCHED_K("ctc_alpha_beta");
  return 0;
}

// backward: dlogits[b,t,v] = scale * (softmax(logits)[v] - occupancy[v]) for t < in_len[b], 0 otherwise (columns
// [V, Vp) zeroed too).  dlogits may alias logits.  `ws` is the workspace filled by cfm_ctc_loss_fwd.
extern "C" int cfm_ctc_loss_bwd(const void* logits, int64_t ld, int B, int T, int V, int Vp, const int* labels, int Lmax,
                                const int* in_len, const int* lab_len, const float* nll, const void* ws, float scale,
                                void* dlogits, int dtype, void* stream) {
  CFM_CHECK_ARG(logits && labels && in_len && lab_len && nll && ws && dlogits, "cfm_ctc_loss_bwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_ctc_loss_bwd: bad dtype");
  CFM_CHECK_ARG(Vp >= V && ld >= Vp, "cfm_ctc_loss_bwd: bad padded vocabulary");
  if (B == 0 || T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int Sp = ((2 * Lmax + 1) + 7) / 8 * 8;
  const float* lse = (const float*)ws;
  const float* lpe = lse + ((size_t)B * T + 63) / 64 * 64;
  const float* alpha = lpe + (size_t)B * T * Sp;
  const float* beta = alpha + (size_t)B * T * Sp;
  const size_t smem = (size_t)V * sizeof(float);
  CFM_CHECK_ARG(smem <= 200 * 1024, "cfm_ctc_loss_bwd: vocabulary %d too large", V);
  if (dtype == CFM_F32) {
    if (smem > 48 * 1024) CFM_SMEM_OPT_IN(ctc_grad_kernel<float>, 200 * 1024);
    ctc_grad_kernel<float><<<B * T, 256, smem, st>>>((const float*)logits, ld, lse, lpe, alpha, beta, nll, labels, Lmax, in_len,
                                                     lab_len, T, Sp, V, Vp, scale, (float*)dlogits);
  } else {
    if (smem > 48 * 1024) CFM_SMEM_OPT_IN(ctc_grad_kernel<__nv_bfloat16>, 200 * 1024);
    ctc_grad_kernel<__nv_bfloat16><<<B * T, 256, smem, st>>>((const __nv_bfloat16*)logits, ld, lse, lpe, alpha, beta, nll, labels,
                                                             Lmax, in_len, lab_len, T, Sp, V, Vp, scale, (__nv_bfloat16*)dlogits);
  }
  CFM_LAUNCHED_K("ctc_grad");
  return 0;
}
