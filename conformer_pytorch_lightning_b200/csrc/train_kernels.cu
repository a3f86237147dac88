// Memory-bound kernels of the TRAINING path (forward pieces that must save state, and every non-GEMM backward):
// LayerNorm fwd (with saved statistics) / bwd, SiLU + dropout fwd / bwd, residual + dropout, GLU fwd / bwd,
// BatchNorm(batch statistics) + SiLU backward, depthwise-conv weight gradient, masked softmax + dropout fwd / bwd,
// column sums (bias gradients).  They implement the autograd of the reference's modules
// (encoder_layer.py:56-70, feedforward.py:16-21, attention.py:88-96, convolution.py:36-48) for the activation dtype
// T in {fp32, bf16}; statistics, the residual stream and every parameter gradient are fp32.
// All are HBM/L2-bandwidth kernels: 128-bit vector accesses, rows walked by whole warps, column sums reduced in
// registers -> shared memory -> one atomicAdd per column and block.
// Dropout is counter based (Philox4x32-10 keyed by (seed, site), counter = element index / 4): the backward kernels
// regenerate the masks, nothing is stored.
#include "cfm_common.cuh"

namespace cfm {
namespace {

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

struct Drop {
  const unsigned long long* seed_dev;   // device scalar: a captured CUDA graph replays with a fresh seed
  uint32_t site;
  uint32_t thr;      // drop when random16 < thr  (thr = p * 2^16)
  float scale;       // 1 / (1 - p); p == 0: thr = 0, scale = 1
};

// multipliers (0 or scale) for VEC consecutive elements starting at element index e (e % VEC == 0, VEC = 4 or 8).
// One Philox4x32-10 call yields 128 bits = the 16-bit randoms of EIGHT elements (element e uses call e / 8, half-word e % 8):
// Philox runs on the integer pipe at half the fp32 rate and, at 32 bits per element, was what bound the dropout kernels
// (~20 integer instructions per element); 16 bits resolve p to 1.5e-5.
template <int VEC>
__device__ __forceinline__ void drop_mult(const Drop& d, unsigned long long e, float (&m)[VEC]) {
  static_assert(VEC == 4 || VEC == 8, "drop_mult: 4 or 8 elements");
  if (d.thr == 0u) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) m[i] = 1.f;
    return;
  }
  const unsigned long long seed = __ldg(d.seed_dev);
  const unsigned long long c = e >> 3;
  const uint4 r = philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)c, (uint32_t)(c >> 32), d.site, 0u);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  const int h0 = VEC == 8 ? 0 : (int)(e & 4);           // first half-word of this call used by element e
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int hw = h0 + i;
    const uint32_t word = (VEC == 8) ? w[hw >> 1] : (h0 ? w[2 + (i >> 1)] : w[i >> 1]);
    const uint32_t r16 = (hw & 1) ? (word >> 16) : (word & 0xffffu);
    m[i] = r16 < d.thr ? 0.f : d.scale;
  }
}

inline Drop make_drop(float p, const uint64_t* seed_dev, int site) {
  Drop d;
  d.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev); d.site = (uint32_t)site;
  if (p <= 0.f || seed_dev == nullptr) { d.thr = 0u; d.scale = 1.f; }
  else {
    double t = (double)p * 65536.0 + 0.5;              // 16-bit threshold: drop when random16 < thr
    d.thr = t >= 65535.0 ? 65535u : (t < 1.0 ? 1u : (uint32_t)t);
    d.scale = (float)(1.0 / (1.0 - (double)d.thr / 65536.0));   // unbiased for the probability actually applied
  }
  return d;
}

// ------------------------------------------------------------------ 16-byte vector access
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                              pack_bf16x2(f[6], f[7]));
  }
};
// load N fp32 values (N = 4 or 8) from an fp32 array
template <int N> __device__ __forceinline__ void load_f32(const float* p, float (&f)[N]) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 v = reinterpret_cast<const float4*>(p)[q];
    f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
  }
}
template <int N> __device__ __forceinline__ void store_f32(float* p, const float (&f)[N]) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) reinterpret_cast<float4*>(p)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
}

template <typename T> __device__ __forceinline__ float act_exp(float x) {       // fp32 path: accurate expf
  if constexpr (sizeof(T) == 4) return expf(x); else return __expf(x);
}
// the value a tensor of type T will hold after the store (bias gradients sum what the GEMMs will read)
template <typename T> __device__ __forceinline__ float rounded(float x) {
  if constexpr (sizeof(T) == 4) return x; else return __bfloat162float(__float2bfloat16_rn(x));
}
template <typename T> __device__ __forceinline__ float dsilu(float h) {       // d/dh [h * sigmoid(h)]
  const float s = act_sigmoid<T>(h);
  return s * fmaf(h, 1.f - s, 1.f);
}

// ------------------------------------------------------------------ (rows x cols) walker with optional column sums
// block = 32 column lanes (VEC columns each) x 8 row lanes; f(row, col0, colsum[VEC]) handles VEC columns of one row.
template <int VEC, typename F>
__device__ __forceinline__ void rowwise(int rows, int cols, float* __restrict__ colsum_out, F&& f) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + tx) * VEC;
  float cs[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) cs[i] = 0.f;
  if (c0 < cols) {
    // four rows per trip: the loads of the later rows are in flight while the first ones are processed (one 16-byte load per
    // thread and trip kept these kernels at ~45 % of the HBM rate)
#pragma unroll 4
    for (int r = blockIdx.y * 8 + ty; r < rows; r += gridDim.y * 8) f(r, c0, cs);
  }
  if (colsum_out != nullptr) {
    __shared__ float red[8][32 * VEC + 1];
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[ty][tx * VEC + i] = cs[i];
    __syncthreads();
    // one 16-byte vector atomic per 4 columns (sm_90+): a quarter of the L2 atomic traffic of scalar adds
    for (int i = threadIdx.x * 4; i < 32 * VEC; i += 256 * 4) {
      const int c = blockIdx.x * 32 * VEC + i;
      if (c < cols) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int y = 0; y < 8; ++y) { s.x += red[y][i]; s.y += red[y][i + 1]; s.z += red[y][i + 2]; s.w += red[y][i + 3]; }
        atomicAdd(reinterpret_cast<float4*>(colsum_out + c), s);
      }
    }
  }
}
inline dim3 rowwise_grid(int rows, int cols, int vec) {
  const int gx = (cols + 32 * vec - 1) / (32 * vec);
  int gy = (rows + 7) / 8;
  const int cap = max(1, 8 * num_sms() / gx);
  if (gy > cap) gy = cap;
  return dim3(gx, max(gy, 1));
}

// ================================================================== LayerNorm
template <int NV, typename TY>
__global__ void __launch_bounds__(256)
ln_fwd_train_kernel(const float* __restrict__ x, int rows, const float* __restrict__ g, const float* __restrict__ b,
                    TY* __restrict__ y, float* __restrict__ mean_o, float* __restrict__ rstd_o,
                    const uint8_t* __restrict__ row_valid, float eps) {
  constexpr int D = NV * 128;
  constexpr float inv_d = 1.0f / D;
  const int lane = threadIdx.x & 31;
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += wpg) {
    float4 v[NV];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) { v[i] = xr[i * 32 + lane]; s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) { mean_o[row] = mean; rstd_o[row] = rstd; }
    const bool keep = (row_valid == nullptr) || (row_valid[row] != 0);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i * 32 + lane);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + i * 32 + lane);
      float4 o;
      o.x = keep ? fmaf(v[i].x * rstd, gg.x, bb.x) : 0.f;
      o.y = keep ? fmaf(v[i].y * rstd, gg.y, bb.y) : 0.f;
      o.z = keep ? fmaf(v[i].z * rstd, gg.z, bb.z) : 0.f;
      o.w = keep ? fmaf(v[i].w * rstd, gg.w, bb.w) : 0.f;
      if constexpr (sizeof(TY) == 4) reinterpret_cast<float4*>(y + (size_t)row * D)[i * 32 + lane] = o;
      else reinterpret_cast<uint2*>(y + (size_t)row * D)[i * 32 + lane] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

// dx_out = dx_in + rstd * (dy*g - mean_d(dy*g) - xhat * mean_d(dy*g*xhat));  dg += dy*xhat;  db += dy   (dy masked by row)
template <int NV, typename TDY>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const TDY* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ g, const uint8_t* __restrict__ row_valid,
              const float* dx_in, float* dx_out, float* __restrict__ dg, float* __restrict__ db, int rows) {
  constexpr int D = NV * 128;
  constexpr float inv_d = 1.0f / D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpg = (gridDim.x * blockDim.x) >> 5;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += wpg) {
    const bool keep = (row_valid == nullptr) || (row_valid[row] != 0);
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], dyv[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 xv = reinterpret_cast<const float4*>(x + (size_t)row * D)[i * 32 + lane];
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      if (!keep) dyv[i] = make_float4(0, 0, 0, 0);
      else if constexpr (sizeof(TDY) == 4) dyv[i] = reinterpret_cast<const float4*>(dy + (size_t)row * D)[i * 32 + lane];
      else {
        const uint2 u = reinterpret_cast<const uint2*>(dy + (size_t)row * D)[i * 32 + lane];
        const float2 a = unpack_bf16x2(u.x), b2 = unpack_bf16x2(u.y);
        dyv[i] = make_float4(a.x, a.y, b2.x, b2.y);
      }
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i * 32 + lane);
      ag[i].x = fmaf(dyv[i].x, xh[i].x, ag[i].x); ag[i].y = fmaf(dyv[i].y, xh[i].y, ag[i].y);
      ag[i].z = fmaf(dyv[i].z, xh[i].z, ag[i].z); ag[i].w = fmaf(dyv[i].w, xh[i].w, ag[i].w);
      ab[i].x += dyv[i].x; ab[i].y += dyv[i].y; ab[i].z += dyv[i].z; ab[i].w += dyv[i].w;
      dyv[i].x *= gg.x; dyv[i].y *= gg.y; dyv[i].z *= gg.z; dyv[i].w *= gg.w;          // dy * g
      c1 += (dyv[i].x + dyv[i].y) + (dyv[i].z + dyv[i].w);
      c2 += (dyv[i].x * xh[i].x + dyv[i].y * xh[i].y) + (dyv[i].z * xh[i].z + dyv[i].w * xh[i].w);
    }
    c1 = warp_sum(c1) * inv_d;
    c2 = warp_sum(c2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = rs * (dyv[i].x - c1 - xh[i].x * c2); o.y = rs * (dyv[i].y - c1 - xh[i].y * c2);
      o.z = rs * (dyv[i].z - c1 - xh[i].z * c2); o.w = rs * (dyv[i].w - c1 - xh[i].w * c2);
      if (dx_in != nullptr) {
        const float4 r = reinterpret_cast<const float4*>(dx_in + (size_t)row * D)[i * 32 + lane];
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      reinterpret_cast<float4*>(dx_out + (size_t)row * D)[i * 32 + lane] = o;
    }
  }
  // block reduction of the parameter gradients, then one atomicAdd per column and block
  __shared__ float4 red[8][32];
  for (int i = 0; i < NV; ++i) {
    for (int which = 0; which < 2; ++which) {
      red[warp][lane] = which == 0 ? ag[i] : ab[i];
      __syncthreads();
      if (warp == 0) {
        float4 s = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) { s.x += red[w][lane].x; s.y += red[w][lane].y; s.z += red[w][lane].z; s.w += red[w][lane].w; }
        atomicAdd(reinterpret_cast<float4*>((which == 0 ? dg : db) + (i * 32 + lane) * 4), s);
      }
      __syncthreads();
    }
  }
}

// ================================================================== SiLU + dropout (feedforward.py:18-19)
template <typename T>
__global__ void __launch_bounds__(256)
silu_dropout_fwd_kernel(const T* __restrict__ h, T* __restrict__ a, int rows, int cols, Drop drop) {
  constexpr int V = Vec<T>::N;
  rowwise<V>(rows, cols, nullptr, [&](int r, int c0, float (&)[V]) {
    const size_t e = (size_t)r * cols + c0;
    float f[V], m[V];
    Vec<T>::load(h + e, f);
    drop_mult<V>(drop, e, m);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = act_silu<T>(f[i]) * m[i];
    Vec<T>::store(a + e, f);
  });
}
// dh = da * mask * silu'(h)  (in place over da allowed);  db += colsum(dh)
template <typename T>
__global__ void __launch_bounds__(256)
silu_dropout_bwd_kernel(const T* da, const T* __restrict__ h, T* dh, float* __restrict__ dbias, int rows, int cols,
                        Drop drop) {
  constexpr int V = Vec<T>::N;
  rowwise<V>(rows, cols, dbias, [&](int r, int c0, float (&cs)[V]) {
    const size_t e = (size_t)r * cols + c0;
    float g[V], f[V], m[V];
    Vec<T>::load(da + e, g);
    Vec<T>::load(h + e, f);
    drop_mult<V>(drop, e, m);
#pragma unroll
    for (int i = 0; i < V; ++i) { g[i] = g[i] * m[i] * dsilu<T>(f[i]); }
    Vec<T>::store(dh + e, g);
    // the bias gradient sums what the GEMMs will see (the stored, rounded values)
#pragma unroll
    for (int i = 0; i < V; ++i) cs[i] += rounded<T>(g[i]);
  });
}

// ================================================================== residual + dropout (encoder_layer.py:58,62,66,69)
// x += alpha * rowmask * dropout(f)
template <typename T>
__global__ void __launch_bounds__(256)
resid_dropout_add_kernel(const float* x_in, float* x_out, const T* __restrict__ f, int rows, int cols, float alpha,
                         const uint8_t* __restrict__ row_valid, Drop drop) {
  constexpr int V = Vec<T>::N;
  rowwise<V>(rows, cols, nullptr, [&](int r, int c0, float (&)[V]) {
    const size_t e = (size_t)r * cols + c0;
    float v[V], m[V], xv[V];
    load_f32<V>(x_in + e, xv);
    if (row_valid == nullptr || row_valid[r] != 0) {
      Vec<T>::load(f + e, v);
      drop_mult<V>(drop, e, m);
#pragma unroll
      for (int i = 0; i < V; ++i) xv[i] = fmaf(alpha * m[i], v[i], xv[i]);
    } else if (x_in == x_out) {
      return;
    }
    store_f32<V>(x_out + e, xv);
  });
}
// df = alpha * rowmask * mask * dx   (fp32 -> T);  dbias += colsum(df)
template <typename T>
__global__ void __launch_bounds__(256)
scale_dropout_bwd_kernel(const float* __restrict__ dx, T* __restrict__ df, float* __restrict__ dbias, int rows, int cols,
                         float alpha, const uint8_t* __restrict__ row_valid, Drop drop) {
  constexpr int V = Vec<T>::N;
  rowwise<V>(rows, cols, dbias, [&](int r, int c0, float (&cs)[V]) {
    const size_t e = (size_t)r * cols + c0;
    float v[V], m[V];
    const bool keep = row_valid == nullptr || row_valid[r] != 0;
    load_f32<V>(dx + e, v);
    drop_mult<V>(drop, e, m);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = keep ? alpha * m[i] * v[i] : 0.f;
    Vec<T>::store(df + e, v);
#pragma unroll
    for (int i = 0; i < V; ++i) cs[i] += rounded<T>(v[i]);
  });
}

// ================================================================== GLU (convolution.py:42)
template <typename T>
__global__ void __launch_bounds__(256)
glu_fwd_kernel(const T* __restrict__ g, T* __restrict__ u, int rows, int d) {
  constexpr int V = Vec<T>::N;
  rowwise<V>(rows, d, nullptr, [&](int r, int c0, float (&)[V]) {
    float a[V], b[V];
    Vec<T>::load(g + (size_t)r * 2 * d + c0, a);
    Vec<T>::load(g + (size_t)r * 2 * d + d + c0, b);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] *= act_sigmoid<T>(b[i]);
    Vec<T>::store(u + (size_t)r * d + c0, a);
  });
}
// dg[:, :d] = du * sig(b);  dg[:, d:] = du * a * sig(b) * (1 - sig(b));  dbias(2d) += colsum(dg)
template <typename T>
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const float* __restrict__ du, const T* __restrict__ g, T* __restrict__ dg, float* __restrict__ dbias,
               int rows, int d) {
  constexpr int V = Vec<T>::N;
  // walk the 2d output columns: the first d are the value half, the rest the gate half
  rowwise<V>(rows, 2 * d, dbias, [&](int r, int c0, float (&cs)[V]) {
    const bool gate = c0 >= d;
    const int c = gate ? c0 - d : c0;
    float a[V], b[V], u[V];
    Vec<T>::load(g + (size_t)r * 2 * d + c, a);
    Vec<T>::load(g + (size_t)r * 2 * d + d + c, b);
    load_f32<V>(du + (size_t)r * d + c, u);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float s = act_sigmoid<T>(b[i]);
      u[i] = gate ? u[i] * a[i] * s * (1.f - s) : u[i] * s;
    }
    Vec<T>::store(dg + (size_t)r * 2 * d + c0, u);
    Vec<T>::load(dg + (size_t)r * 2 * d + c0, u);
#pragma unroll
    for (int i = 0; i < V; ++i) cs[i] += u[i];
  });
}

// ================================================================== BatchNorm(batch stats) + SiLU backward
// n = gamma * xhat + beta, c = silu(n):  dn = dc * silu'(n);  sums[0][ch] += dn * xhat (dgamma), sums[1][ch] += dn (dbeta)
template <typename T>
__global__ void __launch_bounds__(256)
bn_silu_bwd_stats_kernel(const T* __restrict__ dc, const float* __restrict__ raw, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float* __restrict__ sums, int rows, int d) {
  constexpr int V = Vec<T>::N;
  // two column sums per channel: run the walker over 2d "columns" (first d: dgamma, second d: dbeta)
  rowwise<V>(rows, 2 * d, sums, [&](int r, int c0, float (&cs)[V]) {
    const bool second = c0 >= d;
    const int c = second ? c0 - d : c0;
    float g[V], x[V], mu[V], rs[V], ga[V], be[V];
    Vec<T>::load(dc + (size_t)r * d + c, g);
    load_f32<V>(raw + (size_t)r * d + c, x);
    load_f32<V>(mean + c, mu); load_f32<V>(rstd + c, rs); load_f32<V>(gamma + c, ga); load_f32<V>(beta + c, be);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xh = (x[i] - mu[i]) * rs[i];
      const float dn = g[i] * dsilu<T>(fmaf(ga[i], xh, be[i]));
      cs[i] += second ? dn : dn * xh;
    }
  });
}
// draw = gamma * rstd * (dn - dbeta/N - xhat * dgamma/N)
template <typename T>
__global__ void __launch_bounds__(256)
bn_silu_bwd_apply_kernel(const T* __restrict__ dc, const float* __restrict__ raw, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ sums, T* __restrict__ draw, int rows, int d, int batch_stats) {
  constexpr int V = Vec<T>::N;
  // eval-mode BatchNorm (running statistics are constants): no correction terms
  const float inv_n = batch_stats ? 1.f / (float)rows : 0.f;
  rowwise<V>(rows, d, nullptr, [&](int r, int c, float (&)[V]) {
    float g[V], x[V], mu[V], rs[V], ga[V], be[V], sg[V], sb[V];
    Vec<T>::load(dc + (size_t)r * d + c, g);
    load_f32<V>(raw + (size_t)r * d + c, x);
    load_f32<V>(mean + c, mu); load_f32<V>(rstd + c, rs); load_f32<V>(gamma + c, ga); load_f32<V>(beta + c, be);
    load_f32<V>(sums + c, sg); load_f32<V>(sums + d + c, sb);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xh = (x[i] - mu[i]) * rs[i];
      const float dn = g[i] * dsilu<T>(fmaf(ga[i], xh, be[i]));
      g[i] = ga[i] * rs[i] * (dn - sb[i] * inv_n - xh * sg[i] * inv_n);
    }
    Vec<T>::store(draw + (size_t)r * d + c, g);
  });
}

// ================================================================== depthwise conv weight gradient (convolution.py:43)
// dw[j][c] += sum_{b,t} dy[b,t,c] * u[b, t + j - pad, c];  dbias[c] += sum dy.  block = 64 channels x 4 time lanes.
template <typename T, int K>
__global__ void __launch_bounds__(256)
dwconv_wgrad_kernel(const T* __restrict__ dy, const T* __restrict__ u, float* __restrict__ dw, float* __restrict__ dbias,
                    int B, int Tlen, int d, int seg) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int tl = threadIdx.x >> 6;
  const int b = blockIdx.z;
  constexpr int pad = (K - 1) / 2;
  float acc[K + 1];
#pragma unroll
  for (int j = 0; j <= K; ++j) acc[j] = 0.f;
  // each of the 4 time lanes owns SEG/4 = 16 CONSECUTIVE frames of the block's segment: 16 dy values and the 16 + K - 1 u
  // values around them are loaded once into registers (all loads independent, issued back to back) and the K tap sums are
  // fully unrolled FMAs -- the first version re-read u through L1 K times per frame with the loads in the dependent chain
  constexpr int TL = 16;
  const int t0 = blockIdx.y * seg + tl * TL;
  if (c < d && t0 < Tlen) {
    const T* dyb = dy + (size_t)b * Tlen * d + c;
    const T* ub = u + (size_t)b * Tlen * d + c;
    float g[TL], w[TL + K - 1];
#pragma unroll
    for (int i = 0; i < TL; ++i) g[i] = (t0 + i < Tlen) ? to_f32(dyb[(size_t)(t0 + i) * d]) : 0.f;
#pragma unroll
    for (int i = 0; i < TL + K - 1; ++i) {
      const int sidx = t0 + i - pad;
      w[i] = (sidx >= 0 && sidx < Tlen) ? to_f32(ub[(size_t)sidx * d]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < TL; ++i) {
      acc[K] += g[i];
#pragma unroll
      for (int j = 0; j < K; ++j) acc[j] = fmaf(g[i], w[i + j], acc[j]);
    }
  }
  __shared__ float red[4][64];
#pragma unroll 1
  for (int j = 0; j <= K; ++j) {
    red[tl][threadIdx.x & 63] = acc[j];
    __syncthreads();
    if (tl == 0 && c < d) {
      const int i = threadIdx.x & 63;
      const float s = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
      if (j < K) atomicAdd(dw + (size_t)j * d + c, s); else atomicAdd(dbias + c, s);
    }
    __syncthreads();
  }
}

// eight consecutive elements (row offsets are multiples of 8): 16 bytes of bf16 or 2 x 16 bytes of fp32
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) { Vec<__nv_bfloat16>::store(p, f); }
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) { store_f32<8>(p, f); }
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) { Vec<__nv_bfloat16>::load(p, f); }
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) { load_f32<8>(p, f); }

// ================================================================== masked softmax + dropout (attention.py:88-95)
// one warp per (b, h, i) row of S (fp32, row stride Tp); mask semantics of the reference: masked -> -inf, softmax,
// masked -> 0 (a fully masked row gives zeros).  P (and the dropped copy Pd) are written with zeros in columns [T, Tp).
template <typename T>
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(const float* __restrict__ S, T* __restrict__ P, T* __restrict__ Pd, const uint8_t* __restrict__ mask,
                   long long mask_bs, long long mask_rs, int B, int H, int Tq, int Tk, int Tp, Drop drop) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * H * Tq) return;
  const int i = (int)(row % Tq);
  const int b = (int)(row / ((long long)H * Tq));
  const float* s = S + row * Tp;
  const uint8_t* m = mask ? mask + b * mask_bs + i * mask_rs : nullptr;
  if (Tp <= 256) {
    // one pass with the row in registers (8 columns per lane): two 16-byte loads of S, the 8 mask bytes, one 16-byte store
    // of P (and of Pd) per lane, instead of three passes of scalar loads
    const int j0 = lane * 8;
    float v[8];
    bool on[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] = 0.f; on[e] = false; }
    if (j0 < Tp) {
      const float4 a = *reinterpret_cast<const float4*>(s + j0), c4 = *reinterpret_cast<const float4*>(s + j0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) on[e] = (j0 + e < Tk) && (m == nullptr || m[j0 + e] != 0);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < 8; ++e) if (on[e]) mx = fmaxf(mx, v[e]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] = on[e] ? act_exp<T>(v[e] - mx) : 0.f; sum += v[e]; }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (j0 < Tp) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = rounded<T>(v[e] * inv);
      store8(P + row * Tp + j0, v);
      if (Pd != nullptr) {
        float m0[4], m1[4];
        drop_mult<4>(drop, (unsigned long long)row * Tp + j0, m0);
        drop_mult<4>(drop, (unsigned long long)row * Tp + j0 + 4, m1);
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[e] *= m0[e]; v[4 + e] *= m1[e]; }
        store8(Pd + row * Tp + j0, v);
      }
    }
    return;
  }
  float mx = -INFINITY;
  for (int j = lane; j < Tk; j += 32)
    if (m == nullptr || m[j]) mx = fmaxf(mx, s[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Tk; j += 32)
    if (m == nullptr || m[j]) sum += act_exp<T>(s[j] - mx);
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int j0 = lane * 4; j0 < Tp; j0 += 128) {
    float p[4], mult[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + e;
      p[e] = (j < Tk && (m == nullptr || m[j]) && mx > -INFINITY) ? act_exp<T>(s[j] - mx) * inv : 0.f;
    }
    T* po = P + row * Tp + j0;
#pragma unroll
    for (int e = 0; e < 4; ++e) po[e] = from_f32<T>(p[e]);
    if (Pd != nullptr) {
      drop_mult<4>(drop, (unsigned long long)row * Tp + j0, mult);
      T* pd = Pd + row * Tp + j0;
#pragma unroll
      for (int e = 0; e < 4; ++e) pd[e] = from_f32<T>(to_f32(po[e]) * mult[e]);
    }
  }
}
// dS = P * (dP - sum_j dP_j P_j), dP = dPd * dropout multiplier
template <typename T>
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const T* __restrict__ P, const float* __restrict__ dPd, T* __restrict__ dS, long long rows, int Tk,
                   int Tp, Drop drop) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  if (Tp <= 256) {
    const int j0 = lane * 8;
    float pv[8], dp[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { pv[e] = 0.f; dp[e] = 0.f; }
    if (j0 < Tp) {
      float m0[4], m1[4];
      drop_mult<4>(drop, (unsigned long long)row * Tp + j0, m0);
      drop_mult<4>(drop, (unsigned long long)row * Tp + j0 + 4, m1);
      const float4 a = *reinterpret_cast<const float4*>(dPd + row * Tp + j0), c4 = *reinterpret_cast<const float4*>(dPd + row * Tp + j0 + 4);
      const float dd[8] = {a.x * m0[0], a.y * m0[1], a.z * m0[2], a.w * m0[3], c4.x * m1[0], c4.y * m1[1], c4.z * m1[2], c4.w * m1[3]};
      float pl[8];
      load8(P + row * Tp + j0, pl);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (j0 + e < Tk) { pv[e] = pl[e]; dp[e] = dd[e]; }
    }
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) dot = fmaf(pv[e], dp[e], dot);
    dot = warp_sum(dot);
    if (j0 < Tp) {
#pragma unroll
      for (int e = 0; e < 8; ++e) pv[e] *= dp[e] - dot;
      store8(dS + row * Tp + j0, pv);
    }
    return;
  }
  float dot = 0.f;
  for (int j0 = lane * 4; j0 < Tp; j0 += 128) {
    float mult[4];
    drop_mult<4>(drop, (unsigned long long)row * Tp + j0, mult);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (j0 + e < Tk) dot = fmaf(to_f32(P[row * Tp + j0 + e]), dPd[row * Tp + j0 + e] * mult[e], dot);
  }
  dot = warp_sum(dot);
  for (int j0 = lane * 4; j0 < Tp; j0 += 128) {
    float mult[4];
    drop_mult<4>(drop, (unsigned long long)row * Tp + j0, mult);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + e;
      const float v = j < Tk ? to_f32(P[row * Tp + j]) * (dPd[row * Tp + j] * mult[e] - dot) : 0.f;
      dS[row * Tp + j] = from_f32<T>(v);
    }
  }
}

// ================================================================== column sums
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out, int rows, int cols) {
  constexpr int V = Vec<T>::N;
  rowwise<V>(rows, cols, out, [&](int r, int c0, float (&cs)[V]) {
    float v[V];
    Vec<T>::load(x + (size_t)r * ld + c0, v);
#pragma unroll
    for (int i = 0; i < V; ++i) cs[i] += v[i];
  });
}

#define CFM_BY_DTYPE(dtype, CALL)                       \
  do {                                                  \
    if ((dtype) == CFM_F32) { using T = float; CALL; }  \
    else { using T = __nv_bfloat16; CALL; }             \
  } while (0)

}  // namespace
}  // namespace cfm

using namespace cfm;

extern "C" int cfm_ln_fwd_train(const float* x, int rows, int d, const float* g, const float* b, void* y, int y_dtype,
                                float* mean, float* rstd, const uint8_t* row_valid, float eps, void* stream) {
  CFM_CHECK_ARG(x && g && b && y && mean && rstd, "cfm_ln_fwd_train: null pointer");
  CFM_CHECK_ARG(y_dtype == CFM_F32 || y_dtype == CFM_BF16, "cfm_ln_fwd_train: bad dtype");
  CFM_CHECK_ARG(d % 128 == 0 && d >= 128 && d <= 1024, "cfm_ln_fwd_train: d=%d unsupported (need d%%128==0, d<=1024)", d);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = max(1, min((rows + 7) / 8, num_sms() * 8));
#define CFM_LNF(NV)                                                                                                        \
  case NV:                                                                                                                 \
    if (y_dtype == CFM_F32) ln_fwd_train_kernel<NV, float><<<blocks, 256, 0, st>>>(x, rows, g, b, (float*)y, mean, rstd, row_valid, eps); \
    else ln_fwd_train_kernel<NV, __nv_bfloat16><<<blocks, 256, 0, st>>>(x, rows, g, b, (__nv_bfloat16*)y, mean, rstd, row_valid, eps);  \
    break;
  switch (d / 128) { CFM_LNF(1) CFM_LNF(2) CFM_LNF(3) CFM_LNF(4) CFM_LNF(5) CFM_LNF(6) CFM_LNF(7) CFM_LNF(8) }
#undef CFM_LNF
  CFM_LAUNCHED_K("ln_fwd_train");
  return 0;
}

extern "C" int cfm_ln_bwd(const void* dy, int dy_dtype, const float* x, const float* mean, const float* rstd, const float* g,
                          const uint8_t* row_valid, const float* dx_in, float* dx_out, float* dg, float* db, int rows, int d,
                          void* stream) {
  CFM_CHECK_ARG(dy && x && mean && rstd && g && dx_out && dg && db, "cfm_ln_bwd: null pointer");
  CFM_CHECK_ARG(dy_dtype == CFM_F32 || dy_dtype == CFM_BF16, "cfm_ln_bwd: bad dtype");
  CFM_CHECK_ARG(d % 128 == 0 && d >= 128 && d <= 1024, "cfm_ln_bwd: d=%d unsupported (need d%%128==0, d<=1024)", d);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = max(1, min((rows + 31) / 32, num_sms()));      // few blocks: each ends with 2*d/4 vector atomics
#define CFM_LNB(NV)                                                                                                        \
  case NV:                                                                                                                 \
    if (dy_dtype == CFM_F32) ln_bwd_kernel<NV, float><<<blocks, 256, 0, st>>>((const float*)dy, x, mean, rstd, g, row_valid, dx_in, dx_out, dg, db, rows); \
    else ln_bwd_kernel<NV, __nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)dy, x, mean, rstd, g, row_valid, dx_in, dx_out, dg, db, rows);  \
    break;
  switch (d / 128) { CFM_LNB(1) CFM_LNB(2) CFM_LNB(3) CFM_LNB(4) CFM_LNB(5) CFM_LNB(6) CFM_LNB(7) CFM_LNB(8) }
#undef CFM_LNB
  CFM_LAUNCHED_K("ln_bwd");
  return 0;
}

static int check_rc(int rows, int cols, int dtype, const char* who) {
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "%s: bad dtype %d", who, dtype);
  CFM_CHECK_ARG(rows >= 0 && cols > 0 && cols % 8 == 0, "%s: bad shape rows=%d cols=%d (cols %% 8 == 0)", who, rows, cols);
  return 0;
}

extern "C" int cfm_silu_dropout_fwd(const void* h, void* a, int rows, int cols, int dtype, float p, const uint64_t* seed, int site,
                                    void* stream) {
  CFM_CHECK_ARG(h && a, "cfm_silu_dropout_fwd: null pointer");
  if (check_rc(rows, cols, dtype, "cfm_silu_dropout_fwd") != 0) return -1;
  if (rows == 0) return 0;
  const Drop dr = make_drop(p, seed, site);
  CFM_BY_DTYPE(dtype, (silu_dropout_fwd_kernel<T><<<rowwise_grid(rows, cols, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>(
                          (const T*)h, (T*)a, rows, cols, dr)));
  CFM_LAUNCHED_K("silu_dropout_fwd");
  return 0;
}

extern "C" int cfm_silu_dropout_bwd(const void* da, const void* h, void* dh, float* dbias, int rows, int cols, int dtype,
                                    float p, const uint64_t* seed, int site, void* stream) {
  CFM_CHECK_ARG(da && h && dh, "cfm_silu_dropout_bwd: null pointer");
  if (check_rc(rows, cols, dtype, "cfm_silu_dropout_bwd") != 0) return -1;
  if (rows == 0) return 0;
  const Drop dr = make_drop(p, seed, site);
  CFM_BY_DTYPE(dtype, (silu_dropout_bwd_kernel<T><<<rowwise_grid(rows, cols, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>(
                          (const T*)da, (const T*)h, (T*)dh, dbias, rows, cols, dr)));
  CFM_LAUNCHED_K("silu_dropout_bwd");
  return 0;
}

extern "C" int cfm_resid_dropout_add(const float* x_in, float* x, const void* f, int rows, int cols, int dtype, float alpha,
                                     const uint8_t* row_valid, float p, const uint64_t* seed, int site, void* stream) {
  CFM_CHECK_ARG(x && f, "cfm_resid_dropout_add: null pointer");
  if (x_in == nullptr) x_in = x;
  if (check_rc(rows, cols, dtype, "cfm_resid_dropout_add") != 0) return -1;
  if (rows == 0) return 0;
  const Drop dr = make_drop(p, seed, site);
  CFM_BY_DTYPE(dtype, (resid_dropout_add_kernel<T><<<rowwise_grid(rows, cols, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>(
                          x_in, x, (const T*)f, rows, cols, alpha, row_valid, dr)));
  CFM_LAUNCHED_K("resid_dropout_add");
  return 0;
}

extern "C" int cfm_scale_dropout_bwd(const float* dx, void* df, float* dbias, int rows, int cols, int dtype, float alpha,
                                     const uint8_t* row_valid, float p, const uint64_t* seed, int site, void* stream) {
  CFM_CHECK_ARG(dx && df, "cfm_scale_dropout_bwd: null pointer");
  if (check_rc(rows, cols, dtype, "cfm_scale_dropout_bwd") != 0) return -1;
  if (rows == 0) return 0;
  const Drop dr = make_drop(p, seed, site);
  CFM_BY_DTYPE(dtype, (scale_dropout_bwd_kernel<T><<<rowwise_grid(rows, cols, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>(
                          dx, (T*)df, dbias, rows, cols, alpha, row_valid, dr)));
  CFM_LAUNCHED_K("scale_dropout_bwd");
  return 0;
}

extern "C" int cfm_glu_fwd(const void* g, void* u, int rows, int d, int dtype, void* stream) {
  CFM_CHECK_ARG(g && u, "cfm_glu_fwd: null pointer");
  if (check_rc(rows, d, dtype, "cfm_glu_fwd") != 0) return -1;
  if (rows == 0) return 0;
  CFM_BY_DTYPE(dtype, (glu_fwd_kernel<T><<<rowwise_grid(rows, d, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>((const T*)g, (T*)u,
                                                                                                              rows, d)));
  CFM_LAUNCHED_K("glu_fwd");
  return 0;
}

extern "C" int cfm_glu_bwd(const float* du, const void* g, void* dg, float* dbias, int rows, int d, int dtype, void* stream) {
  CFM_CHECK_ARG(du && g && dg, "cfm_glu_bwd: null pointer");
  if (check_rc(rows, d, dtype, "cfm_glu_bwd") != 0) return -1;
  if (rows == 0) return 0;
  CFM_BY_DTYPE(dtype, (glu_bwd_kernel<T><<<rowwise_grid(rows, 2 * d, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>(
                          du, (const T*)g, (T*)dg, dbias, rows, d)));
  CFM_LAUNCHED_K("glu_bwd");
  return 0;
}

extern "C" int cfm_bn_silu_bwd(const void* dc, const float* raw, const float* mean, const float* rstd, const float* gamma,
                               const float* beta, float* sums, void* draw, int rows, int d, int dtype, int batch_stats,
                               void* stream) {
  CFM_CHECK_ARG(dc && raw && mean && rstd && gamma && beta && sums && draw, "cfm_bn_silu_bwd: null pointer");
  if (check_rc(rows, d, dtype, "cfm_bn_silu_bwd") != 0) return -1;
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CFM_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * d, st));
  CFM_BY_DTYPE(dtype, (bn_silu_bwd_stats_kernel<T><<<rowwise_grid(rows, 2 * d, Vec<T>::N), 256, 0, st>>>(
                          (const T*)dc, raw, mean, rstd, gamma, beta, sums, rows, d)));
  CFM_LAUNCHED_K("bn_silu_bwd_stats");
  CFM_BY_DTYPE(dtype, (bn_silu_bwd_apply_kernel<T><<<rowwise_grid(rows, d, Vec<T>::N), 256, 0, st>>>(
                          (const T*)dc, raw, mean, rstd, gamma, beta, sums, (T*)draw, rows, d, batch_stats)));
  CFM_LAUNCHED_K("bn_silu_bwd_apply");
  return 0;
}

extern "C" int cfm_dwconv_wgrad(const void* dy, const void* u, float* dw, float* dbias, int B, int Tlen, int d, int k,
                                int dtype, void* stream) {
  CFM_CHECK_ARG(dy && u && dw && dbias, "cfm_dwconv_wgrad: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_dwconv_wgrad: bad dtype");
  CFM_CHECK_ARG(k == 7 || k == 15 || k == 31, "cfm_dwconv_wgrad: kernel size %d unsupported (7, 15, 31)", k);
  CFM_CHECK_ARG(B >= 0 && Tlen >= 0 && B <= 65535, "cfm_dwconv_wgrad: bad B/T");
  if (B == 0 || Tlen == 0) return 0;
  const int seg = 64;
  dim3 grid((d + 63) / 64, (Tlen + seg - 1) / seg, B);
  cudaStream_t st = (cudaStream_t)stream;
#define CFM_DWG(KK) CFM_BY_DTYPE(dtype, (dwconv_wgrad_kernel<T, KK><<<grid, 256, 0, st>>>((const T*)dy, (const T*)u, dw, dbias, B, Tlen, d, seg)))
  if (k == 7) CFM_DWG(7); else if (k == 15) CFM_DWG(15); else CFM_DWG(31);
#undef CFM_DWG
  CFM_LAUNCHED_K("dwconv_wgrad");
  return 0;
}

extern "C" int cfm_softmax_fwd(const float* S, void* P, void* Pd, const uint8_t* mask, int64_t mask_bs, int64_t mask_rs, int B,
                               int H, int Tq, int Tk, int Tp, int dtype, float p, const uint64_t* seed, int site, void* stream) {
  CFM_CHECK_ARG(S && P, "cfm_softmax_fwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_softmax_fwd: bad dtype");
  CFM_CHECK_ARG(Tp >= Tk && Tp % 8 == 0, "cfm_softmax_fwd: row stride %d must be a multiple of 8 and >= Tk=%d", Tp, Tk);
  CFM_CHECK_ARG(p <= 0.f || Pd != nullptr, "cfm_softmax_fwd: dropout needs the Pd output");
  const long long rows = (long long)B * H * Tq;
  if (rows == 0) return 0;
  const Drop dr = make_drop(p, seed, site);
  const int blocks = (int)((rows + 7) / 8);
  CFM_BY_DTYPE(dtype, (softmax_fwd_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(S, (T*)P, p > 0.f ? (T*)Pd : nullptr, mask,
                                                                                        mask_bs, mask_rs, B, H, Tq, Tk, Tp, dr)));
  CFM_LAUNCHED_K("softmax_fwd");
  return 0;
}

extern "C" int cfm_softmax_bwd(const void* P, const float* dPd, void* dS, int B, int H, int Tq, int Tk, int Tp, int dtype,
                               float p, const uint64_t* seed, int site, void* stream) {
  CFM_CHECK_ARG(P && dPd && dS, "cfm_softmax_bwd: null pointer");
  CFM_CHECK_ARG(dtype == CFM_F32 || dtype == CFM_BF16, "cfm_softmax_bwd: bad dtype");
  CFM_CHECK_ARG(Tp >= Tk && Tp % 8 == 0, "cfm_softmax_bwd: bad row stride");
  const long long rows = (long long)B * H * Tq;
  if (rows == 0) return 0;
  const Drop dr = make_drop(p, seed, site);
  const int blocks = (int)((rows + 7) / 8);
  CFM_BY_DTYPE(dtype, (softmax_bwd_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>((const T*)P, dPd, (T*)dS, rows, Tk, Tp, dr)));
  CFM_LAUNCHED_K("softmax_bwd");
  return 0;
}

extern "C" int cfm_colsum(const void* x, int64_t ld, float* out, int rows, int cols, int dtype, void* stream) {
  CFM_CHECK_ARG(x && out, "cfm_colsum: null pointer");
  if (check_rc(rows, cols, dtype, "cfm_colsum") != 0) return -1;
  CFM_CHECK_ARG(ld >= cols && ld % 8 == 0, "cfm_colsum: bad leading dimension");
  if (rows == 0) return 0;
  CFM_BY_DTYPE(dtype, (colsum_kernel<T><<<rowwise_grid(rows, cols, Vec<T>::N), 256, 0, (cudaStream_t)stream>>>((const T*)x, ld, out,
                                                                                                                rows, cols)));
  CFM_LAUNCHED_K("colsum");
  return 0;
}

// ------------------------------------------------------------------ optimizer step (module.py:140-143: torch.optim.Adam)
namespace cfm {
namespace {
// One Adam update over a flat fp32 segment (torch.optim.Adam semantics, no weight decay / amsgrad):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// optionally writing the bf16 copy of the new parameter values the compute path consumes.  HBM bound: 28 B per element.
template <bool VEC>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, __nv_bfloat16* __restrict__ p16, int64_t n, int head, float step_size,
                                                   float b1, float omb1, float b2, float omb2, float eps, float inv_sqrt_bc2,
                                                   float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {      // omb = 1 - beta, rounded from double like PyTorch's
    gg *= gscale;
    mm = fmaf(b1, mm, omb1 * gg);
    vv = fmaf(b2, vv, omb2 * gg * gg);
    pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
  };
  if (VEC) {
    // the four fp32 pointers share their 16-byte phase: `head` scalar elements, 16-byte vectors, scalar tail
    const int64_t n4 = (n - head) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + head);
    float4* m4 = reinterpret_cast<float4*>(m + head);
    float4* v4 = reinterpret_cast<float4*>(v + head);
    const float4* g4 = reinterpret_cast<const float4*>(g + head);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      const float4 gg = __ldg(g4 + i);
      upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
      if (p16) reinterpret_cast<uint2*>(p16 + head)[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
    }
    const int64_t tail0 = head + (n4 << 2);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < head + (n - tail0); k += stride) {
      const int64_t i = k < head ? k : tail0 + (k - head);
      upd(p[i], g[i], m[i], v[i]);
      if (p16) p16[i] = __float2bfloat16_rn(p[i]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      upd(p[i], g[i], m[i], v[i]);
      if (p16) p16[i] = __float2bfloat16_rn(p[i]);
    }
  }
}
}  // namespace
}  // namespace cfm

extern "C" int cfm_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, double lr, double beta1,
                             double beta2, double eps, int step, float grad_scale, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(p && g && m && v, "cfm_adam_step: null pointer");
  CFM_CHECK_ARG(n >= 0 && step >= 1, "cfm_adam_step: bad size / step");
  CFM_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "cfm_adam_step: bad hyper-parameters");
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
  const float step_size = (float)(lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float b1f = (float)beta1, omb1 = (float)(1.0 - beta1), b2f = (float)beta2, omb2 = (float)(1.0 - beta2), epsf = (float)eps;
  const uintptr_t ph = reinterpret_cast<uintptr_t>(p) & 15;
  int head = (int)(((16 - ph) & 15) / 4);
  if (head > n) head = (int)n;
  const bool vec = (ph & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == ph && (reinterpret_cast<uintptr_t>(m) & 15) == ph &&
                   (reinterpret_cast<uintptr_t>(v) & 15) == ph &&
                   (p_bf16 == nullptr || ((reinterpret_cast<uintptr_t>(p_bf16) + 2 * (uintptr_t)head) & 7) == 0);
  const int64_t work = vec ? (n + 3) / 4 + 8 : n;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (vec)
    adam_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n, head, step_size, b1f,
                                                                         omb1, b2f, omb2, epsf, inv_sqrt_bc2, grad_scale);
  else
    adam_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (__nv_bfloat16*)p_bf16, n, 0, step_size, b1f,
                                                                          omb1, b2f, omb2, epsf, inv_sqrt_bc2, grad_scale);
  CFM_LAUNCHED_K("adam");
  return 0;
}
