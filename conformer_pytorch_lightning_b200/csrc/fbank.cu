// Feature front-end on device (scope row f1): Kaldi-compatible log-mel filterbank + global CMVN, i.e. what the reference
// computes on the CPU with torchaudio.compliance.kaldi.fbank (processor.py:185-191) and GlobalCMVN (cmvn.py:22-33).
//   fbank_frames   warp per frame: snip-edges framing (25 ms / 10 ms), DC removal, pre-emphasis 0.97 (first sample
//                  replicated), povey window -> (frames, 400) fp32
//   (DFT)          the 512-point real DFT of a 400-sample frame is a GEMM with a (514 x 400) [cos | -sin] basis --
//                  cfm_gemm_ex, fp32 accumulation -- there is no FFT butterfly pass and no 512-padded copy
//   fbank_power    |X|^2 of the 257 bins -> (frames, 264) fp32 (padded row stride for 16-byte vector access)
//   (mel)          (frames, 257) x (80, 257)^T -- cfm_gemm_ex again
//   fbank_log_cmvn log(max(mel, eps)) (+ (x - mean) * istd), frames past an utterance's end behave like the reference's
//                  zero padding of the feature matrix (pad_sequence at processor.py:302-304, then CMVN in the encoder)
//   cmvn           stand-alone (x - mean) * istd for features that arrive already computed
// All memory bound, fp32 throughout (the parity gate is 1e-4 on the features).
#include "cfm_common.cuh"
#include <float.h>

namespace cfm {
namespace {

constexpr int WIN = 400, SHIFT = 160;

__global__ void __launch_bounds__(256)
fbank_frames_kernel(const float* __restrict__ wave, long long wave_bs, const int* __restrict__ n_samples,
                    const float* __restrict__ window, float* __restrict__ frames, int B, int m_max, float preemph) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= (long long)B * m_max) return;
  const int b = (int)(row / m_max), i = (int)(row % m_max);
  const int n = n_samples[b];
  const int m = n < WIN ? 0 : 1 + (n - WIN) / SHIFT;
  float* out = frames + row * WIN;
  if (i >= m) {
    for (int j = lane; j < WIN; j += 32) out[j] = 0.f;
    return;
  }
  const float* w = wave + b * wave_bs + (long long)i * SHIFT;
  float s = 0.f;
  for (int j = lane; j < WIN; j += 32) s += w[j];
  const float mean = warp_sum(s) * (1.0f / WIN);
  for (int j = lane; j < WIN; j += 32) {
    const float x = w[j] - mean;
    const float xp = w[j > 0 ? j - 1 : 0] - mean;
    out[j] = (x - preemph * xp) * window[j];
  }
}

__global__ void __launch_bounds__(256)
fbank_power_kernel(const float* __restrict__ spec, int ld_spec, float* __restrict__ power, int ld_pow, long long rows, int bins) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld_pow) return;
  const long long r = i / ld_pow;
  const int k = (int)(i % ld_pow);
  float v = 0.f;
  if (k < bins) {
    const float re = spec[r * ld_spec + k], im = spec[r * ld_spec + bins + k];
    v = re * re + im * im;
  }
  power[i] = v;
}

__global__ void __launch_bounds__(256)
fbank_log_cmvn_kernel(const float* __restrict__ mel, float* __restrict__ out, const int* __restrict__ n_samples,
                      const float* __restrict__ mean, const float* __restrict__ istd, int B, int m_max, int nmel) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * m_max * nmel) return;
  const int c = (int)(i % nmel);
  const long long row = i / nmel;
  const int b = (int)(row / m_max), fr = (int)(row % m_max);
  const int n = n_samples[b];
  const int m = n < WIN ? 0 : 1 + (n - WIN) / SHIFT;
  float v = fr < m ? logf(fmaxf(mel[i], FLT_EPSILON)) : 0.f;
  if (mean != nullptr) v -= mean[c];
  if (istd != nullptr) v *= istd[c];
  out[i] = v;
}

__global__ void __launch_bounds__(256)
cmvn_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ istd,
            long long n, int d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % d);
  float v = x[i] - mean[c];
  if (istd != nullptr) v *= istd[c];
  y[i] = v;
}

}  // namespace
}  // namespace cfm

using namespace cfm;

extern "C" int cfm_fbank_frames(const float* wave, int64_t wave_bs, const int* n_samples, const float* window, float* frames,
                                int B, int m_max, float preemph, void* stream) {
  CFM_CHECK_ARG(wave && n_samples && window && frames, "cfm_fbank_frames: null pointer");
  if (B <= 0 || m_max <= 0) return 0;
  const long long rows = (long long)B * m_max;
  fbank_frames_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(wave, wave_bs, n_samples, window, frames, B,
                                                                                   m_max, preemph);
  CFM_LAUNCHED_K("fbank_frames");
  return 0;
}

extern "C" int cfm_fbank_power(const float* spec, int ld_spec, float* power, int ld_pow, int64_t rows, int bins, void* stream) {
  CFM_CHECK_ARG(spec && power && ld_spec >= 2 * bins && ld_pow >= bins, "cfm_fbank_power: bad arguments");
  if (rows <= 0) return 0;
  const long long n = rows * ld_pow;
  fbank_power_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(spec, ld_spec, power, ld_pow, rows, bins);
  CFM_LAUNCHED_K("fbank_power");
  return 0;
}

extern "C" int cfm_fbank_log_cmvn(const float* mel, float* out, const int* n_samples, const float* mean, const float* istd, int B,
                                  int m_max, int nmel, void* stream) {
  CFM_CHECK_ARG(mel && out && n_samples, "cfm_fbank_log_cmvn: null pointer");
  if (B <= 0 || m_max <= 0) return 0;
  const long long n = (long long)B * m_max * nmel;
  fbank_log_cmvn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mel, out, n_samples, mean, istd, B, m_max, nmel);
  CFM_LAUNCHED_K("fbank_log_cmvn");
  return 0;
}

extern "C" int cfm_cmvn(const float* x, float* y, const float* mean, const float* istd, int64_t n, int d, void* stream) {
  CFM_CHECK_ARG(x && y && mean && d > 0, "cfm_cmvn: bad arguments");
  if (n <= 0) return 0;
  cmvn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, mean, istd, n, d);
  CFM_LAUNCHED_K("cmvn");
  return 0;
}
