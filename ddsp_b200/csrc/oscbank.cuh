// core.oscillator_bank (core.py:911-962) as a stand-alone op on audio-rate
// envelopes [B, N, K] (what synths.Sinusoidal feeds it).
//   amp = 0 where f >= sr/2; omega = f * 2 pi / sr; phi = cumsum_t(omega);
//   out = amp * sin(phi)  ([B,N,K]) or its sum over k ([B,N]).
// The cumsum is an exact wrapping sum of 64-bit fixed-point turns, done as a
// three-pass chunked scan (chunk = 128 samples): per-chunk totals, a scan of the
// totals per (b, k), then the running phase inside each chunk.  Threads run over
// k (the contiguous axis), so every global access is coalesced; the path is
// HBM-bound: f is read twice, amp once.
#pragma once
#include "common.cuh"

namespace ddsp {

constexpr int kObChunk = 128;
constexpr int kObThreads = 128;

// pass 1: chunk totals.  grid (n_chunks, B), threads over k.
__global__ void __launch_bounds__(kObThreads)
oscbank_chunk_sums(const float* __restrict__ f, unsigned long long* __restrict__ sums,
                   int N, int K, int n_chunks, double inv_sr) {
  const int b = blockIdx.y, ch = blockIdx.x;
  const int t0 = ch * kObChunk, t1 = min(N, t0 + kObChunk);
  for (int k = threadIdx.x; k < K; k += kObThreads) {
    const float* fp = f + ((size_t)b * N + t0) * K + k;
    unsigned long long acc = 0;
    for (int t = t0; t < t1; ++t, fp += K) acc += turns_to_fix64((double)(*fp) * inv_sr);
    sums[((size_t)b * n_chunks + ch) * K + k] = acc;
  }
}

// pass 2: exclusive scan of the chunk totals along the chunk axis, in place.
__global__ void __launch_bounds__(kObThreads)
oscbank_scan_chunks(unsigned long long* __restrict__ sums, int K, int n_chunks,
                    int64_t BK) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BK) return;
  const int64_t b = i / K;
  const int k = (int)(i - b * K);
  unsigned long long run = 0;
  unsigned long long* p = sums + (size_t)b * n_chunks * K + k;
  for (int ch = 0; ch < n_chunks; ++ch, p += K) {
    const unsigned long long v = *p;
    *p = run;
    run += v;
  }
}

// pass 3: running phase inside the chunk, sin, mask, optional sum over k.
template <bool SUM>
__global__ void __launch_bounds__(kObThreads)
oscbank_apply(const float* __restrict__ f, const float* __restrict__ a,
              const unsigned long long* __restrict__ offs, float* __restrict__ out,
              int N, int K, int n_chunks, double inv_sr, float nyquist) {
  __shared__ float partial[kObChunk][kObThreads / 32 + 1];
  const int b = blockIdx.y, ch = blockIdx.x;
  const int t0 = ch * kObChunk, t1 = min(N, t0 + kObChunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (SUM) {
    for (int i = threadIdx.x; i < kObChunk * (kObThreads / 32 + 1); i += kObThreads)
      (&partial[0][0])[i] = 0.f;
    __syncthreads();
  }
  for (int kb = 0; kb < K; kb += kObThreads) {
    const int k = kb + threadIdx.x;
    const bool live = k < K;
    unsigned long long ph = live ? offs[((size_t)b * n_chunks + ch) * K + k] : 0ull;
    const size_t base = ((size_t)b * N + t0) * K + (live ? k : 0);
    const float* fp = f + base;
    const float* ap = a + base;
    for (int t = t0; t < t1; ++t, fp += K, ap += K) {
      float v = 0.f;
      if (live) {
        const float fv = *fp;
        ph += turns_to_fix64((double)fv * inv_sr);
        const uint32_t p32 = (uint32_t)((ph + 0x80000000ull) >> 32);
        const float s = sinpif((float)(int)p32 * 4.656612873077393e-10f);
        v = (fv >= nyquist) ? 0.f : (*ap) * s;
        if (!SUM) out[((size_t)b * N + t) * K + k] = v;
      }
      if (SUM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) partial[t - t0][warp] += v;
      }
    }
  }
  if (SUM) {
    __syncthreads();
    for (int i = threadIdx.x; i < t1 - t0; i += kObThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kObThreads / 32; ++w) s += partial[i][w];
      out[(size_t)b * N + t0 + i] = s;
    }
  }
}

}  // namespace ddsp
