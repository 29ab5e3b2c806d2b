// core.fft_convolve for ONE long impulse response per item (the LTI case of
// core.py:1382-1473: effects.Reverb, effects.py:103-117 - 48000 taps on 64000
// samples).  The direct form costs N * S MACs (3e9 per item), so this one really is
// a frequency-domain convolution - hand-written: uniformly partitioned
// overlap-save with L = 1024-sample blocks and 2048-point complex FFTs in shared
// memory (the reference takes three 131072-point FFTs per item through cuFFT).
//
//   * REAL -> COMPLEX PACKING WITHOUT WASTE.  Convolution with a real h is linear
//     over complex inputs, so the two time-halves of an item ride in one complex
//     signal: z[n] = x[n] + i x[n + N2], w = z * h, y[n] = Re w[n] + Im w[n - N2].
//   * NO REORDERING PASS.  The forward transform is decimation-in-frequency (natural
//     in, digit-reversed out), the inverse decimation-in-time (digit-reversed in,
//     natural out); the spectra only ever meet in element-wise products, so both
//     sides simply live in digit-reversed order.  Radix 4 (2048 = 4^5 * 2).
//   * Overlap-save: input window j = samples [(j-1) L, (j+1) L) of z; IR partition
//     p = h[p L, (p+1) L) zero-padded to 2 L; output block j = last L samples of
//     IFFT(sum_p Z_{j-p} H_p).
// Kernels: lc_fft_blocks (windows of z, or partitions of h -> spectra),
// lc_mac_ifft (spectral multiply-accumulate over the partitions + inverse FFT ->
// w blocks; four output blocks per CTA, windows sliding through registers; only
// the blocks the crop needs are produced), lc_combine (Re / Im recombination, delay
// crop, optional accumulate).  The backward pass (d audio, d impulse response) is
// the same three kernels on time-reversed operands (`reverse` flag).
#pragma once
#include "common.cuh"

namespace ddsp {
namespace lc {

constexpr int L = 1024;          // block length
constexpr int M = 2 * L;         // FFT size
constexpr int LOGM = 11;
constexpr int THREADS = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {   // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}

constexpr int TWN = 3 * M / 4;      // twiddles tw[m] = exp(-2 pi i m / M), m < 3 M / 4

__device__ __forceinline__ void fill_twiddles(float2* tw, int tid) {
  for (int m = tid; m < TWN; m += THREADS) {
    float s, c;
    sincospif(-2.0f * (float)m / (float)M, &s, &c);
    tw[m] = make_float2(c, s);
  }
}

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)
__device__ __forceinline__ float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }   // a * (+i)

// Forward transform, decimation in frequency, 2048 = 4^5 * 2: five radix-4 stages
// (quarter spans 512, 128, 32, 8, 2) and one radix-2 stage; natural order in,
// digit-reversed order out.  A stage is y = T F4 x per group (F4 the 4-point DFT, T
// the twiddles W^(k m)); six barriers instead of the eleven of a radix-2 network.
__device__ __forceinline__ void fft_dif(float2* s, const float2* tw, int tid) {
#pragma unroll 1
  for (int q = M / 4, sh = 0; q >= 2; q >>= 2, sh += 2) {      // W = tw[1 << sh]
    __syncthreads();
    for (int i = tid; i < M / 4; i += THREADS) {
      const int k = i & (q - 1);
      const int base = ((i - k) << 2) + k;
      const float2 x0 = s[base], x1 = s[base + q], x2 = s[base + 2 * q], x3 = s[base + 3 * q];
      const float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
      const int t = k << sh;
      s[base] = cadd(a, c);
      s[base + q] = cmul(cadd(b, mul_mi(d)), tw[t]);
      s[base + 2 * q] = cmul(csub(a, c), tw[2 * t]);
      s[base + 3 * q] = cmul(cadd(b, mul_pi(d)), tw[3 * t]);
    }
  }
  __syncthreads();
  for (int i = tid; i < M / 2; i += THREADS) {                 // radix-2, span 1
    const float2 u = s[2 * i], v = s[2 * i + 1];
    s[2 * i] = cadd(u, v);
    s[2 * i + 1] = csub(u, v);
  }
  __syncthreads();
}

// Inverse (unscaled): the stages of fft_dif undone in reverse order, each as
// x = F4^H conj(T) y; digit-reversed order in, natural order out.
__device__ __forceinline__ void ifft_dit(float2* s, const float2* tw, int tid) {
  __syncthreads();
  for (int i = tid; i < M / 2; i += THREADS) {
    const float2 u = s[2 * i], v = s[2 * i + 1];
    s[2 * i] = cadd(u, v);
    s[2 * i + 1] = csub(u, v);
  }
#pragma unroll 1
  for (int q = 2, sh = 8; q <= M / 4; q <<= 2, sh -= 2) {
    __syncthreads();
    for (int i = tid; i < M / 4; i += THREADS) {
      const int k = i & (q - 1);
      const int base = ((i - k) << 2) + k;
      const int t = k << sh;
      const float2 x0 = s[base], x1 = cmul_conj(s[base + q], tw[t]),
                   x2 = cmul_conj(s[base + 2 * q], tw[2 * t]),
                   x3 = cmul_conj(s[base + 3 * q], tw[3 * t]);
      const float2 a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
      s[base] = cadd(a, c);
      s[base + q] = cadd(b, mul_pi(d));
      s[base + 2 * q] = csub(a, c);
      s[base + 3 * q] = cadd(b, mul_mi(d));
    }
  }
  __syncthreads();
}

// mode 0: window j of z (two time-halves of audio item b packed as re / im);
// mode 1: partition p of the impulse response of item b (real, zero-padded).
// reverse: the source is read back to front (x[len - 1 - n]) - the backward pass
// convolves with time-reversed signals.  grid (n_blocks, items).
__global__ void __launch_bounds__(THREADS)
lc_fft_blocks(const float* __restrict__ src, float2* __restrict__ spec, int len,
              int n2, int n_blocks, int mode, int reverse) {
  __shared__ float2 s[M];
  __shared__ float2 tw[TWN];
  const int tid = threadIdx.x, j = blockIdx.x, b = blockIdx.y;
  const float* x = src + (size_t)b * len;
  fill_twiddles(tw, tid);
  for (int i = tid; i < M; i += THREADS) {
    float re = 0.f, im = 0.f;
    if (mode == 0) {
      const long long n = (long long)(j - 1) * L + i;          // index into z
      if (n >= 0 && n < n2) {
        re = x[reverse ? len - 1 - n : n];
        if (n + n2 < len) im = x[reverse ? len - 1 - (n + n2) : n + n2];
      }
    } else if (i < L) {
      const long long n = (long long)j * L + i;
      if (n < len) re = x[reverse ? len - 1 - n : n];
    }
    s[i] = make_float2(re, im);
  }
  fft_dif(s, tw, tid);
  float2* out = spec + ((size_t)b * n_blocks + j) * M;
  for (int i = tid; i < M; i += THREADS) out[i] = s[i];
}

// w blocks j0 .. j0 + JT - 1 of item b: block j = last L samples of
// IFFT(sum_p Z[b, j - p] H[bi, p]).  One CTA owns JT consecutive output blocks and
// walks the partitions once: per partition ONE spectrum of H and ONE new window of
// Z are loaded (the other JT - 1 windows slide through registers), so the L2
// traffic per output block is 2 P / JT spectra instead of 2 P.
// grid (ceil(n_blocks / JT), B); blocks j_first + [0, n_blocks) are produced.
// Z: [B, n_in, M], H: [Bi, P, M], W: [B, n_out * L].
constexpr int JT = 4;
constexpr int EPT = 4;                  // spectrum elements per thread per pass
constexpr int PASSES = M / (THREADS * EPT);
constexpr size_t kMacSmem = sizeof(float2) * ((size_t)JT * M + TWN);

__global__ void __launch_bounds__(THREADS, 2)
lc_mac_ifft(const float2* __restrict__ Z, const float2* __restrict__ H,
            float2* __restrict__ W, int n_in, int P, int n_out, int ir_batch_stride,
            int j_first, int n_blocks) {
  extern __shared__ __align__(16) unsigned char lc_smem[];
  float2* sAcc = reinterpret_cast<float2*>(lc_smem);            // [JT][M]
  float2* tw = sAcc + (size_t)JT * M;                           // [TWN]
  const int tid = threadIdx.x, b = blockIdx.y;
  const int j0 = j_first + blockIdx.x * JT;
  const int jt_n = min(JT, j_first + n_blocks - j0);
  fill_twiddles(tw, tid);
  const float2* Zb = Z + (size_t)b * n_in * M;
  const float2* Hb = H + (size_t)b * ir_batch_stride;
  // partitions that reach any block of the tile: window j0 + k - p in [0, n_in)
  const int p_lo = max(0, j0 - (n_in - 1)), p_hi = min(P - 1, j0 + JT - 1);
#pragma unroll 1
  for (int pass = 0; pass < PASSES; ++pass) {       // the spectrum in two halves
    const int f0 = pass * THREADS * EPT + tid;
    float2 acc[JT][EPT], zw[JT][EPT];
#pragma unroll
    for (int k = 0; k < JT; ++k)
#pragma unroll
      for (int e = 0; e < EPT; ++e) acc[k][e] = make_float2(0.f, 0.f);
    auto load_window = [&](int w, float2 (&dst)[EPT]) {
      if (w >= 0 && w < n_in) {
        const float2* zp = Zb + (size_t)w * M + f0;
#pragma unroll
        for (int e = 0; e < EPT; ++e) dst[e] = zp[e * THREADS];
      } else {
#pragma unroll
        for (int e = 0; e < EPT; ++e) dst[e] = make_float2(0.f, 0.f);
      }
    };
    // slot convention: at partition p = p_lo + u (u mod JT), window j0 + k - p sits
    // in zw[(k - u) & (JT - 1)]
#pragma unroll
    for (int k = 0; k < JT; ++k) load_window(j0 + k - p_lo, zw[k]);
    for (int pb = p_lo; pb <= p_hi; pb += JT) {
#pragma unroll
      for (int u = 0; u < JT; ++u) {
        const int p = pb + u;
        if (p <= p_hi) {
          const float2* hp = Hb + (size_t)p * M + f0;
          float2 h[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) h[e] = hp[e * THREADS];
#pragma unroll
          for (int k = 0; k < JT; ++k) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
              const float2 z = zw[(k - u) & (JT - 1)][e];
              acc[k][e].x = fmaf(z.x, h[e].x, fmaf(-z.y, h[e].y, acc[k][e].x));
              acc[k][e].y = fmaf(z.x, h[e].y, fmaf(z.y, h[e].x, acc[k][e].y));
            }
          }
          // next partition: windows shift down by one; the slot of the highest
          // window (k = JT - 1) is refilled with window j0 - (p + 1)
          load_window(j0 - (p + 1), zw[(JT - 1 - u) & (JT - 1)]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < JT; ++k)
#pragma unroll
      for (int e = 0; e < EPT; ++e) sAcc[(size_t)k * M + f0 + e * THREADS] = acc[k][e];
  }
  const float scale = 1.0f / (float)M;
  for (int k = 0; k < jt_n; ++k) {
    float2* s = sAcc + (size_t)k * M;
    ifft_dit(s, tw, tid);
    float2* out = W + ((size_t)b * n_out + (j0 + k)) * L;
    for (int i = tid; i < L; i += THREADS) {
      const float2 v = s[L + i];
      out[i] = make_float2(v.x * scale, v.y * scale);
    }
  }
}

// y[b, n] = Re w[n + start] + Im w[n + start - N2], n < out_len (crop of
// crop_and_compensate_delay, core.py:1338-1379).  w is valid on [0, n_out * L).
__global__ void __launch_bounds__(256)
lc_combine(const float2* __restrict__ W, float* __restrict__ out, int n2, int w_len,
           int start, int out_len, int total_len, int accumulate, int w_lo, int w_hi) {
  const int b = blockIdx.y;
  const float2* w = W + (size_t)b * w_len;
  float* o = out + (size_t)b * out_len;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < out_len;
       n += gridDim.x * blockDim.x) {
    const int t = n + start;                   // position in the full convolution
    float v = 0.f;
    if (t < total_len) {
      // only [w_lo, w_hi) was produced; w is identically zero outside the blocks
      // that any partition reaches
      if (t >= w_lo && t < w_hi) v = w[t].x;
      const int u = t - n2;
      if (u >= w_lo && u < w_hi) v += w[u].y;
    }
    if (accumulate) v += o[n];
    o[n] = v;
  }
}

struct Geom {
  int n2, n_in, P, n_out, w_len;
};
__host__ inline Geom geom(int N, int S) {
  Geom g;
  g.n2 = (N + 1) / 2;
  g.n_in = (g.n2 + L - 1) / L + 1;             // windows with any non-zero sample
  g.P = (S + L - 1) / L;
  // w = z * h has n2 + S - 1 samples
  g.n_out = (g.n2 + S - 1 + L - 1) / L;
  g.w_len = g.n_out * L;
  return g;
}

}  // namespace lc
}  // namespace ddsp
