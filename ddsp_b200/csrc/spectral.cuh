// Pieces of the multi-scale spectrogram loss (ddsp/losses.py:130-243 with
// spectral_ops.stft, spectral_ops.py:34-47) that torch runs as ~70 elementwise
// launches per FFT size: framing + periodic-Hann windowing (and its adjoint, the
// windowed overlap-add of frame gradients), and the L1 magnitude / log-magnitude
// differences with their gradient w.r.t. the complex STFT.  The FFTs themselves
// stay cuFFT (torch.fft.rfft and its autograd) as SURVEY 8f-1 prescribes.
// All three kernels are plain streaming kernels: 128-bit accesses, one pass.
#pragma once
#include "common.cuh"

namespace ddsp {

// frames[b, t, i] = window[i] * audio[b, t * step + i]   (0 beyond N: pad_end=True)
__global__ void __launch_bounds__(256)
frame_window_kernel(const float* __restrict__ audio, const float* __restrict__ window,
                    float* __restrict__ frames, int N, int T, int n, int step) {
  const int b = blockIdx.y;
  const long long e = 4ll * ((long long)blockIdx.x * 256 + threadIdx.x);
  if (e >= (long long)T * n) return;
  const int t = (int)(e / n), i = (int)(e - (long long)t * n);
  const float* a = audio + (size_t)b * N;
  const long long s = (long long)t * step + i;
  const float4 w = *reinterpret_cast<const float4*>(window + i);
  float4 v;
  v.x = (s + 0 < N) ? a[s + 0] * w.x : 0.f;
  v.y = (s + 1 < N) ? a[s + 1] * w.y : 0.f;
  v.z = (s + 2 < N) ? a[s + 2] * w.z : 0.f;
  v.w = (s + 3 < N) ? a[s + 3] * w.w : 0.f;
  *reinterpret_cast<float4*>(frames + ((size_t)b * T) * n + e) = v;
}

// grad_audio[b, s] (+)= scale * sum_t window[s - t step] * grad_frames[b, t, s - t step]
// scale_ptr: optional DEVICE scalar (the upstream gradient of the loss value), so
// the six FFT sizes of the multi-scale loss accumulate into one buffer without a
// host round trip or an elementwise pass.
__global__ void __launch_bounds__(256)
frame_window_adjoint_kernel(const float* __restrict__ gframes,
                            const float* __restrict__ window,
                            float* __restrict__ gaudio, int N, int T, int n, int step,
                            const float* __restrict__ scale_ptr, int accumulate) {
  const int b = blockIdx.y;
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= N) return;
  const float* g = gframes + ((size_t)b * T) * n;
  int t_hi = s / step;
  if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = (s - n + step) / step;          // ceil((s - n + 1) / step) for s-n+1 > 0
  if (s - n + 1 <= 0) t_lo = 0;
  float acc = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int i = s - t * step;
    if (i >= 0 && i < n) acc = fmaf(window[i], g[(size_t)t * n + i], acc);
  }
  if (scale_ptr != nullptr) acc *= *scale_ptr;
  float* o = gaudio + (size_t)b * N + s;
  if (accumulate) acc += *o;
  *o = acc;
}

// One pass over the two complex STFTs: sums of |mag_t - mag_a| and
// |safe_log mag_t - safe_log mag_a| (core.safe_log, core.py:213-216), and the
// gradient of  w_mag * mean|.| + w_log * mean|.|  w.r.t. X_a (PyTorch's complex
// convention: dL/dRe + i dL/dIm).
__global__ void __launch_bounds__(256)
spectral_l1_kernel(const float2* __restrict__ xt, const float2* __restrict__ xa,
                   float2* __restrict__ grad, double* __restrict__ sums, long long M,
                   float w_mag, float w_log, float inv_count, float eps,
                   int n_bins, int irfft_scale) {
  double s_mag = 0.0, s_log = 0.0;
  const long long stride = (long long)gridDim.x * 256 * 2;
  for (long long e = 2 * ((long long)blockIdx.x * 256 + threadIdx.x); e < M; e += stride) {
    float4 T4, A4;
    const bool two = e + 1 < M;
    if (two) {
      T4 = *reinterpret_cast<const float4*>(xt + e);
      A4 = *reinterpret_cast<const float4*>(xa + e);
    } else {
      const float2 t1 = xt[e], a1 = xa[e];
      T4 = make_float4(t1.x, t1.y, 0.f, 0.f);
      A4 = make_float4(a1.x, a1.y, 0.f, 0.f);
    }
    float g[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float tr = h ? T4.z : T4.x, ti = h ? T4.w : T4.y;
      const float ar = h ? A4.z : A4.x, ai = h ? A4.w : A4.y;
      const float mt = sqrtf(tr * tr + ti * ti), ma = sqrtf(ar * ar + ai * ai);
      const float d1 = mt - ma;
      const float lt = logf(mt <= 0.f ? eps : mt), la = logf(ma <= 0.f ? eps : ma);
      const float d2 = lt - la;
      if (h == 0 || two) {
        s_mag += fabsf(d1);
        s_log += fabsf(d2);
      }
      const float sg1 = (d1 > 0.f) ? 1.f : (d1 < 0.f ? -1.f : 0.f);
      const float sg2 = (d2 > 0.f) ? 1.f : (d2 < 0.f ? -1.f : 0.f);
      float dma = 0.f;                         // dL / d mag_a
      if (ma > 0.f) {
        const float inv = 1.0f / ma;
        dma = -(w_mag * sg1 + w_log * sg2 * inv) * inv_count * inv;   // times X_a / ma
      }
      if (irfft_scale) {
        // pre-scale for the transpose of rfft written as a plain irfft:
        // d/dx = n * irfft(Y), Y = G at DC / Nyquist, G / 2 in between
        // irfft_scale = -1: the caller's inverse transform is unnormalised
        // (norm='forward'), so only the halving of the interior bins remains
        const int k = (int)((e + h) % n_bins);
        const float sc = irfft_scale < 0 ? 1.0f : (float)irfft_scale;
        dma *= (k == 0 || k == n_bins - 1) ? sc : 0.5f * sc;
      }
      g[2 * h] = dma * ar;
      g[2 * h + 1] = dma * ai;
    }
    if (two) *reinterpret_cast<float4*>(grad + e) = make_float4(g[0], g[1], g[2], g[3]);
    else grad[e] = make_float2(g[0], g[1]);
  }
  // block reduce, one double atomic per sum per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_mag += __shfl_xor_sync(0xffffffffu, s_mag, o);
    s_log += __shfl_xor_sync(0xffffffffu, s_log, o);
  }
  __shared__ double sh[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s_mag; sh[1][warp] = s_log; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; c += sh[1][w]; }
    atomicAdd(sums, a);
    atomicAdd(sums + 1, c);
  }
}

}  // namespace ddsp
