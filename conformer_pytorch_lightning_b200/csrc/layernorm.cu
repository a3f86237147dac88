// LayerNorm (+ optional chained second LayerNorm + optional row mask), one warp per row.
// HBM-bound: reads the fp32 residual row once with 128-bit loads, keeps it in registers for
// the two-pass statistics, writes the normalised row once (bf16 or fp32).
// Replaces nn.LayerNorm at encoder_layer.py:56,59,63,67,70 and encoder.py:74 of the reference.
#include "cfm_common.cuh"

namespace cfm {
namespace {

template <int NV>  // NV float4 per lane -> d = NV * 128
__device__ __forceinline__ void ln_row(float4 (&v)[NV], const float* __restrict__ g,
                                       const float* __restrict__ b, int lane, float eps) {
  constexpr float inv_d = 1.0f / (NV * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i * 32 + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + i * 32 + lane);
    v[i].x = fmaf(v[i].x * rstd, gg.x, bb.x);
    v[i].y = fmaf(v[i].y * rstd, gg.y, bb.y);
    v[i].z = fmaf(v[i].z * rstd, gg.z, bb.z);
    v[i].w = fmaf(v[i].w * rstd, gg.w, bb.w);
  }
}

template <int NV, typename TY>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* x, int rows, const float* __restrict__ g1,
                 const float* __restrict__ b1, float* x_out,  // x_out may alias x
                 const float* __restrict__ g2, const float* __restrict__ b2, TY* __restrict__ y,
                 const uint8_t* __restrict__ row_valid, float eps) {
  constexpr int D = NV * 128;
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps_per_grid) {
    float4 v[NV];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
    ln_row<NV>(v, g1, b1, lane, eps);
    if (x_out != nullptr) {
      float4* xo = reinterpret_cast<float4*>(x_out + (size_t)row * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) xo[i * 32 + lane] = v[i];
    }
    if (g2 != nullptr) ln_row<NV>(v, g2, b2, lane, eps);
    if (y != nullptr) {
      const bool keep = (row_valid == nullptr) || (row_valid[row] != 0);
      if (!keep) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if constexpr (sizeof(TY) == 4) {
        float4* yo = reinterpret_cast<float4*>(y + (size_t)row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i) yo[i * 32 + lane] = v[i];
      } else {
        uint2* yo = reinterpret_cast<uint2*>(y + (size_t)row * D);
#pragma unroll
        for (int i = 0; i < NV; ++i)
          yo[i * 32 + lane] = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
      }
    }
  }
}

template <int NV>
int launch_ln(const float* x, int rows, const float* g1, const float* b1, float* x_out, const float* g2,
              const float* b2, void* y, int y_dtype, const uint8_t* rv, float eps, cudaStream_t st) {
  const int blocks = max(1, min((rows + 7) / 8, num_sms() * 8));
  if (y_dtype == CFM_F32)
    CFM_CUDA_OK(launch_pdl(layernorm_kernel<NV, float>, dim3(blocks), dim3(256), 0, st, 1, x, rows, g1, b1, x_out, g2, b2,
                           (float*)y, rv, eps));
  else
    CFM_CUDA_OK(launch_pdl(layernorm_kernel<NV, __nv_bfloat16>, dim3(blocks), dim3(256), 0, st, 1, x, rows, g1, b1, x_out,
                           g2, b2, (__nv_bfloat16*)y, rv, eps));
  CFM_LAUNCHED_K("layernorm");
  return 0;
}

}  // namespace
}  // namespace cfm

extern "C" int cfm_layernorm(const float* x, int rows, int d, const float* g1, const float* b1,
                             float* x_out, const float* g2, const float* b2, void* y, int y_dtype,
                             const uint8_t* row_valid, float eps, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(x && g1 && b1, "cfm_layernorm: null x/g1/b1");
  CFM_CHECK_ARG((g2 == nullptr) == (b2 == nullptr), "cfm_layernorm: g2/b2 must both be set or both null");
  CFM_CHECK_ARG(y_dtype == CFM_F32 || y_dtype == CFM_BF16, "cfm_layernorm: bad y_dtype %d", y_dtype);
  CFM_CHECK_ARG(d % 128 == 0 && d >= 128 && d <= 1024, "cfm_layernorm: d=%d unsupported (need d%%128==0, d<=1024)", d);
  CFM_CHECK_ARG(rows >= 0, "cfm_layernorm: rows<0");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d / 128) {
    case 1: return launch_ln<1>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    case 2: return launch_ln<2>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    case 3: return launch_ln<3>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    case 4: return launch_ln<4>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    case 5: return launch_ln<5>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    case 6: return launch_ln<6>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    case 7: return launch_ln<7>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
    default: return launch_ln<8>(x, rows, g1, b1, x_out, g2, b2, y, y_dtype, row_valid, eps, st);
  }
}

// ------------------------------------------------------------------ L2 prefetch of the next layer's weights
// Every kernel of the stack is a single wave that leaves 8..24 SMs idle (124-140 CTAs on 148 SMs) and starts by pulling
// its weights from HBM (they are cold: ~5.6 MB per layer, first touch of the step).  A few CTAs of this kernel run on the
// idle SMs of the PREVIOUS layer's kernels (side stream, a parallel branch of the CUDA graph) and pull the next layer's
// weights into L2, so the weight rings of the fused kernels see L2 latency instead of HBM latency.
namespace cfm {
namespace {
__global__ void __launch_bounds__(256)
l2_prefetch_kernel(const uint8_t* __restrict__ p, size_t bytes) {
  const size_t lines = (bytes + 127) / 128;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p + i * 128));
}
}  // namespace
}  // namespace cfm

namespace cfm {
namespace {
struct PrefetchList { const uint8_t* p[8]; unsigned long long bytes[8]; int n; };
__global__ void __launch_bounds__(256)
l2_prefetch_multi_kernel(const PrefetchList pl) {
  for (int t = 0; t < pl.n; ++t) {
    const size_t lines = (pl.bytes[t] + 127) / 128;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (size_t)gridDim.x * blockDim.x)
      asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pl.p[t] + i * 128));
  }
}
}  // namespace
}  // namespace cfm

extern "C" int cfm_l2_prefetch_multi(const void* const* ptrs, const int64_t* bytes, int n, int blocks, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(ptrs != nullptr && bytes != nullptr && n >= 0 && n <= 8, "cfm_l2_prefetch_multi: up to 8 regions");
  if (n == 0) return 0;
  PrefetchList pl{};
  pl.n = n;
  for (int i = 0; i < n; ++i) { pl.p[i] = (const uint8_t*)ptrs[i]; pl.bytes[i] = (unsigned long long)bytes[i]; }
  if (blocks <= 0) blocks = 16;
  l2_prefetch_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pl);
  CFM_LAUNCHED_K("l2_prefetch");
  return 0;
}

extern "C" int cfm_l2_prefetch(const void* p, int64_t bytes, int blocks, void* stream) {
  using namespace cfm;
  CFM_CHECK_ARG(p != nullptr && bytes >= 0, "cfm_l2_prefetch: bad arguments");
  if (bytes == 0) return 0;
  if (blocks <= 0) blocks = 16;
  l2_prefetch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)p, (size_t)bytes);
  CFM_LAUNCHED_K("l2_prefetch");
  return 0;
}
