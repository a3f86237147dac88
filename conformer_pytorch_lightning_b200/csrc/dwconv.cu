// Memory-bound kernels of the Conformer conv module and small helpers:
//   * depthwise conv1d along time (k = 15 / 31) + folded BatchNorm affine + SiLU on channel-last
//     activations, shared-memory halo staged along time, 128-bit coalesced loads
//     (replaces depthwise_conv -> norm -> activation, convolution.py:43-45 of the reference);
//   * BatchNorm batch statistics / apply for the training path (convolution.py:44);
//   * the streaming relative-position fold (attention.py:78-88, P == Tk).
#include "cfm_common.cuh"

namespace cfm {
namespace {

constexpr int kDwTT = 64;   // output frames per block
constexpr int kDwCG = 64;   // channels per block (one 128-byte line of bf16)
constexpr int kDwR = 16;    // outputs per thread along time (4 warps x 16 = 64)

// stage rows [t0 - pad, t0 + TT + pad) x 64 channels as fp32 in shared memory; zero outside [0,T).
// All 16-byte global loads of a thread are issued back to back (MAXL independent loads in flight) before any
// of them is converted and stored: the kernel is latency bound otherwise (ncu: long-scoreboard stalls on the
// first use of each load with ~1 wave of CTAs).
template <typename T, int MAXL>
__device__ __forceinline__ void dw_stage(const T* __restrict__ xb, float* smem, int t0, int Tlen, int d,
                                         int c0, int rows_needed, int pad) {
  constexpr int V = Act<T>::kVec;            // elements per 16-byte load
  constexpr int CHUNKS = kDwCG / V;          // 16-byte chunks per staged row
  const int total = rows_needed * CHUNKS;
  uint4 raw[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    const int idx = threadIdx.x + l * 128;
    const int r = idx / CHUNKS, ch = idx % CHUNKS;
    const int t = t0 - pad + r;
    raw[l] = make_uint4(0u, 0u, 0u, 0u);
    if (idx < total && t >= 0 && t < Tlen)
      raw[l] = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)t * d + c0 + ch * V));
  }
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    const int idx = threadIdx.x + l * 128;
    if (idx >= total) break;
    const int r = idx / CHUNKS, ch = idx % CHUNKS;
    float* dst = smem + r * kDwCG + ch * V;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<uint4*>(dst) = raw[l];
    } else {
      const float2 a = unpack_bf16x2(raw[l].x), b = unpack_bf16x2(raw[l].y);
      const float2 c = unpack_bf16x2(raw[l].z), e = unpack_bf16x2(raw[l].w);
      *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(c.x, c.y, e.x, e.y);
    }
  }
}

// K > 0: compile-time taps (fully unrolled register blocking); K == 0: runtime taps (generic path)
template <typename T, int K, bool SILU>
__global__ void __launch_bounds__(128)
dwconv_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              void* __restrict__ y, int Tlen, int d, int k_rt) {
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int k = (K > 0) ? K : k_rt;
  const int pad = (k - 1) / 2;
  const int t0 = blockIdx.x * kDwTT;
  const int c0 = blockIdx.y * kDwCG;
  const int b = blockIdx.z;
  const T* xb = x + (size_t)b * Tlen * d;
  const int rows_needed = kDwTT + k - 1;
  // loads per thread: (64 + k - 1) rows x (64 channels / elements-per-16-bytes) chunks over 128 threads
  constexpr int KMAX = (K > 0) ? K : 31;
  constexpr int MAXL = ((kDwTT + KMAX - 1) * (kDwCG / Act<T>::kVec) + 127) / 128;
  dw_stage<T, MAXL>(xb, smem, t0, Tlen, d, c0, rows_needed, pad);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + 2 * lane;
  const float2 bb = *reinterpret_cast<const float2*>(bias + c);
  float2 acc[kDwR];
#pragma unroll
  for (int r = 0; r < kDwR; ++r) acc[r] = bb;
  const float* sm = smem + (warp * kDwR) * kDwCG + 2 * lane;

  if constexpr (K > 0) {
    float2 wr[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wr[j] = __ldg(reinterpret_cast<const float2*>(w + (size_t)j * d + c));
#pragma unroll
    for (int s = 0; s < kDwR + K - 1; ++s) {
      const float2 xv = *reinterpret_cast<const float2*>(sm + s * kDwCG);
#pragma unroll
      for (int r = 0; r < kDwR; ++r) {
        const int j = s - r;
        if (j >= 0 && j < K) {
          acc[r].x = fmaf(xv.x, wr[j].x, acc[r].x);
          acc[r].y = fmaf(xv.y, wr[j].y, acc[r].y);
        }
      }
    }
  } else {
    for (int j = 0; j < k; ++j) {
      const float2 wj = __ldg(reinterpret_cast<const float2*>(w + (size_t)j * d + c));
#pragma unroll
      for (int r = 0; r < kDwR; ++r) {
        const float2 xv = *reinterpret_cast<const float2*>(sm + (r + j) * kDwCG);
        acc[r].x = fmaf(xv.x, wj.x, acc[r].x);
        acc[r].y = fmaf(xv.y, wj.y, acc[r].y);
      }
    }
  }

#pragma unroll
  for (int r = 0; r < kDwR; ++r) {
    const int t = t0 + warp * kDwR + r;
    if (t < Tlen) {
      const size_t off = ((size_t)b * Tlen + t) * d + c;
      if constexpr (SILU) {
        const float v0 = act_silu<T>(acc[r].x), v1 = act_silu<T>(acc[r].y);
        if constexpr (sizeof(T) == 4)
          *reinterpret_cast<float2*>((float*)y + off) = make_float2(v0, v1);
        else
          *reinterpret_cast<uint32_t*>((__nv_bfloat16*)y + off) = pack_bf16x2(v0, v1);
      } else {
        *reinterpret_cast<float2*>((float*)y + off) = acc[r];   // raw conv + bias, fp32 (training path)
      }
    }
  }
}

template <typename T, bool SILU>
int launch_dw(const void* x, const float* w, const float* bias, void* y, int B, int Tlen, int d, int k,
              cudaStream_t st) {
  dim3 grid((Tlen + kDwTT - 1) / kDwTT, d / kDwCG, B);
  const size_t smem = (size_t)(kDwTT + k - 1) * kDwCG * sizeof(float);
  const T* xx = (const T*)x;
  if (k == 15) CFM_CUDA_OK(launch_pdl(dwconv_kernel<T, 15, SILU>, grid, dim3(128), smem, st, 1, xx, w, bias, y, Tlen, d, k));
  else if (k == 31) CFM_CUDA_OK(launch_pdl(dwconv_kernel<T, 31, SILU>, grid, dim3(128), smem, st, 1, xx, w, bias, y, Tlen, d, k));
  else CFM_CUDA_OK(launch_pdl(dwconv_kernel<T, 0, SILU>, grid, dim3(128), smem, st, 1, xx, w, bias, y, Tlen, d, k));
  CFM_LAUNCHED_K("dwconv");
  return 0;
}

// ------------------------------------------------------------------ BatchNorm (training) pieces
// Deterministic (no atomics: the training forward repeats bit for bit under the same seed): one block owns 16 channels and
// reduces ALL rows itself -- 64 row lanes per channel, fixed-order tree in shared memory.  d / 16 blocks (16 at d = 256)
// stream 256 KB each; the first version used 64-channel slabs + atomicAdd (order dependent sums, and 15.7 us at the C5 shard
// because only 64 blocks were launched).
__global__ void __launch_bounds__(1024)
bn_stats_kernel(const float* __restrict__ x, int rows, int d, float* __restrict__ sum,
                float* __restrict__ sumsq) {
  const int cl = threadIdx.x & 15, rl = threadIdx.x >> 4;       // 16 channels x 64 row lanes
  const int c = blockIdx.x * 16 + cl;
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  if (c < d) {
    int r = rl;
    for (; r + 64 < rows; r += 128) {                             // two independent chains / loads in flight
      const float a = x[(size_t)r * d + c], b = x[(size_t)(r + 64) * d + c];
      s0 += a; q0 = fmaf(a, a, q0);
      s1 += b; q1 = fmaf(b, b, q1);
    }
    if (r < rows) { const float a = x[(size_t)r * d + c]; s0 += a; q0 = fmaf(a, a, q0); }
  }
  __shared__ float ss[64][17], sq[64][17];
  ss[rl][cl] = s0 + s1;
  sq[rl][cl] = q0 + q1;
  __syncthreads();
#pragma unroll
  for (int w = 32; w >= 1; w >>= 1) {
    if (rl < w) { ss[rl][cl] += ss[rl + w][cl]; sq[rl][cl] += sq[rl + w][cl]; }
    __syncthreads();
  }
  if (rl == 0 && c < d) { sum[c] += ss[0][cl]; sumsq[c] += sq[0][cl]; }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_silu_kernel(const float* __restrict__ x, size_t n4, int d, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ beta, T* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i * 4) % d);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 m = *reinterpret_cast<const float4*>(mean + c), r = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
    const float o0 = act_silu<T>(fmaf((v.x - m.x) * r.x, g.x, be.x));
    const float o1 = act_silu<T>(fmaf((v.y - m.y) * r.y, g.y, be.y));
    const float o2 = act_silu<T>(fmaf((v.z - m.z) * r.z, g.z, be.z));
    const float o3 = act_silu<T>(fmaf((v.w - m.w) * r.w, g.w, be.w));
    if constexpr (sizeof(T) == 4)
      reinterpret_cast<float4*>(y)[i] = make_float4(o0, o1, o2, o3);
    else
      reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
  }
}

// ------------------------------------------------------------------ streaming rel-pos fold
template <typename T>
__global__ void __launch_bounds__(128)
relpos_keys_kernel(const T* __restrict__ k, int64_t k_bs, int64_t k_ts, const T* __restrict__ p, int64_t p_bs,
                   const float* __restrict__ u, const float* __restrict__ vb, T* __restrict__ k_out,
                   float* __restrict__ key_bias, int B, int H, int Tk) {
  // one warp per (j, h); lane owns channels 2*lane, 2*lane+1 of the 64-wide head
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= Tk * H) return;
  const int j = wid / H, h = wid % H;
  const int c = h * 64 + 2 * lane;
  const float d0 = vb[c] - u[c], d1 = vb[c + 1] - u[c + 1];
  for (int b = 0; b < B; ++b) {
    const T* pr = p + b * p_bs + (size_t)j * H * 64 + c;
    const float p0 = to_f32(pr[0]), p1 = to_f32(pr[1]);
    const float kb = warp_sum(fmaf(d0, p0, d1 * p1));
    const T* kr = k + b * k_bs + j * k_ts + c;
    T* ko = k_out + ((size_t)b * Tk + j) * H * 64 + c;
    ko[0] = from_f32<T>(to_f32(kr[0]) + p0);
    ko[1] = from_f32<T>(to_f32(kr[1]) + p1);
    if (lane == 0) key_bias[((size_t)b * H + h) * Tk + j] = kb;
  }
}

}  // namespace
}  // namespace cfm

extern "C" int cfm_dwconv(const void* x, const float* w, const float* bias, void* y, int B, int T, int d,
                          int k, int dtype, int apply_silu, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(x && w && bias && y, "cfm_dwconv: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_dwconv: bad dtype %d", dtype);
  CFM_CHECK_ARG(d % kDwCG == 0, "cfm_dwconv: d=%d must be a multiple of %d", d, kDwCG);
  CFM_CHECK_ARG(k >= 1 && k <= 31 && (k & 1), "cfm_dwconv: kernel size %d unsupported (odd, <= 31)", k);
  CFM_CHECK_ARG(B >= 0 && T >= 0 && B <= 65535, "cfm_dwconv: bad B/T");
  if (B == 0 || T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == CFM_F32)
    return apply_silu ? launch_dw<float, true>(x, w, bias, y, B, T, d, k, st)
                      : launch_dw<float, false>(x, w, bias, y, B, T, d, k, st);
  return apply_silu ? launch_dw<__nv_bfloat16, true>(x, w, bias, y, B, T, d, k, st)
                    : launch_dw<__nv_bfloat16, false>(x, w, bias, y, B, T, d, k, st);
}

extern "C" int cfm_bn_stats(const float* x, int rows, int d, float* sum, float* sumsq, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(x && sum && sumsq, "cfm_bn_stats: null pointer");
  if (rows <= 0) return 0;
  bn_stats_kernel<<<(d + 15) / 16, 1024, 0, (cudaStream_t)stream>>>(x, rows, d, sum, sumsq);
  CFM_LAUNCHED_K("bn_stats");
  return 0;
}

extern "C" int cfm_bn_apply_silu(const float* x, int rows, int d, const float* mean, const float* rstd,
                                 const float* gamma, const float* beta, void* y, int dtype, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(x && mean && rstd && gamma && beta && y, "cfm_bn_apply_silu: null pointer");
  CFM_CHECK_ARG(d % 4 == 0, "cfm_bn_apply_silu: d %% 4 != 0");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_bn_apply_silu: bad dtype");
  if (rows <= 0) return 0;
  const size_t n4 = (size_t)rows * d / 4;
  const int blocks = (int)min((n4 + 255) / 256, (size_t)num_sms() * 8);
  if (dtype == CFM_F32)
    bn_apply_silu_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n4, d, mean, rstd, gamma, beta, (float*)y);
  else
    bn_apply_silu_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n4, d, mean, rstd, gamma, beta,
                                                                                (__nv_bfloat16*)y);
  CFM_LAUNCHED_K("bn_apply_silu");
  return 0;
}

extern "C" int cfm_relpos_keys(const void* k, int64_t k_bs, int64_t k_ts, const void* p, int64_t p_bs, const float* u,
                               const float* vb, void* k_out, float* key_bias, int B, int H, int Tk,
                               int dtype, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(k && p && u && vb && k_out && key_bias, "cfm_relpos_keys: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_relpos_keys: bad dtype");
  if (B <= 0 || Tk <= 0) return 0;
  const int warps = Tk * H;
  const int blocks = (warps + 3) / 4;
  if (dtype == CFM_F32)
    relpos_keys_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>((const float*)k, k_bs, k_ts, (const float*)p, p_bs, u, vb,
                                                                     (float*)k_out, key_bias, B, H, Tk);
  else
    relpos_keys_kernel<__nv_bfloat16><<<blocks, 128, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)k, k_bs, k_ts, (const __nv_bfloat16*)p, p_bs, u, vb, (__nv_bfloat16*)k_out, key_bias, B, H, Tk);
  CFM_LAUNCHED_K("relpos_keys");
  return 0;
}
