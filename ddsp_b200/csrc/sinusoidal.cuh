// Frame-rate oscillator bank with PER-SINUSOID frequencies - the fused form of
//   resample(frequencies, N) + resample(amplitudes, N, amp_method) +
//   core.oscillator_bank(...)                      (core.py:573-642, 911-962)
// for synths.Sinusoidal.get_signal (synths.py:305-323) and for
// core.harmonic_synthesis with harmonic_shifts (core.py:1084-1093, where
// f_k = f0 * k * (1 + shift_k) is no longer an integer multiple of one phase).
// The [B, N, K] envelopes the reference materialises never exist.
//
// Per sinusoid k the frequency is piecewise linear in time (v1 bilinear, frame
// F := frame F-1), so inside frame i the inclusive phase sum has the same closed
// form as the harmonic kernel, with per-k tables:
//   phi_k(i*hop + r) = P_ik + (r+1) a_ik + (a_{i+1,k} - a_ik)/hop * r(r+1)/2,
//   a = f/sr in turns, P_ik = sum_{j<i} [hop a_jk + (a_{j+1,k} - a_jk)(hop-1)/2]
// kept as 64-bit fixed-point turns (wrapping adds are exact).  Three passes:
//   1. per (b, tile, k) total of the frame sums         (grid n_tiles x B)
//   2. exclusive scan over the tiles per (b, k)         (oscbank_scan_chunks)
//   3. tables P/A/D for the tile in shared memory, then one thread per sample
//      loops over k: phase -> sinpif, Nyquist mask on the reference's float32
//      envelope (lo + (hi - lo) * frac, no FMA), two-row amplitude interpolation.
#pragma once
#include "common.cuh"
#include "oscbank.cuh"

namespace ddsp {

constexpr int kSfThreads = 128;

__device__ __forceinline__ unsigned long long sf_frame_total(float f_lo, float f_hi,
                                                             int hop, double inv_sr) {
  const double a0 = (double)f_lo * inv_sr, a1 = (double)f_hi * inv_sr;
  return turns_to_fix64((double)hop * a0 + (a1 - a0) * (0.5 * (hop - 1)));
}

// pass 1.  sums[b, tile, k] = sum of the frame totals of the tile's frames.
__global__ void __launch_bounds__(kSfThreads)
sinus_tile_sums(const float* __restrict__ f, unsigned long long* __restrict__ sums,
                int F, int K, int hop, int FT, int n_tiles, double inv_sr) {
  const int b = blockIdx.y, tile = blockIdx.x;
  const int i0 = tile * FT, i1 = min(F, i0 + FT);
  for (int k = threadIdx.x; k < K; k += kSfThreads) {
    const float* fp = f + ((size_t)b * F) * K + k;
    unsigned long long acc = 0;
    float cur = fp[(size_t)i0 * K];
    for (int i = i0; i < i1; ++i) {
      const float nxt = fp[(size_t)min(i + 1, F - 1) * K];
      acc += sf_frame_total(cur, nxt, hop, inv_sr);
      cur = nxt;
    }
    sums[((size_t)b * n_tiles + tile) * K + k] = acc;
  }
}

struct SfSmem {
  size_t off_P, off_A, off_D, off_f, off_a, total;
};
__host__ __device__ inline SfSmem sf_smem(int FT, int K) {
  SfSmem s;
  size_t o = 0;
  s.off_P = o; o += sizeof(unsigned long long) * (size_t)FT * K;
  s.off_A = o; o += sizeof(unsigned long long) * (size_t)FT * K;
  s.off_D = o; o += sizeof(unsigned long long) * (size_t)FT * K;
  s.off_f = o; o += sizeof(float) * (size_t)(FT + 1) * K;
  s.off_a = o; o += sizeof(float) * (size_t)(FT + 1) * K;
  s.total = o;
  return s;
}

// pass 3.
template <bool WINDOW>
__global__ void __launch_bounds__(kSfThreads)
sinus_apply(const float* __restrict__ f, const float* __restrict__ a,
            const unsigned long long* __restrict__ offs, float* __restrict__ out,
            int F, int K, int N, int hop, int FT, int n_tiles, double inv_sr,
            float nyquist, int accumulate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SfSmem L = sf_smem(FT, K);
  unsigned long long* sP = (unsigned long long*)(smem_raw + L.off_P);
  unsigned long long* sA = (unsigned long long*)(smem_raw + L.off_A);
  unsigned long long* sD = (unsigned long long*)(smem_raw + L.off_D);
  float* sF = (float*)(smem_raw + L.off_f);
  float* sAm = (float*)(smem_raw + L.off_a);
  const int b = blockIdx.y, tile = blockIdx.x;
  const int i0 = tile * FT, nfr = min(FT, F - i0);
  const int tid = threadIdx.x;

  // rows i0 .. i0 + nfr (frame F := frame F-1)
  for (int idx = tid; idx < (nfr + 1) * K; idx += kSfThreads) {
    const int r = idx / K, k = idx - r * K;
    const size_t g = ((size_t)b * F + min(i0 + r, F - 1)) * K + k;
    sF[idx] = f[g];
    sAm[idx] = a[g];
  }
  __syncthreads();
  // per-k tables: a running wrapping sum over the tile's frames
  for (int k = tid; k < K; k += kSfThreads) {
    unsigned long long P = offs[((size_t)b * n_tiles + tile) * K + k];
    for (int j = 0; j < nfr; ++j) {
      const float f_lo = sF[j * K + k], f_hi = sF[(j + 1) * K + k];
      const double a0 = (double)f_lo * inv_sr, a1 = (double)f_hi * inv_sr;
      sP[j * K + k] = P + 0x80000000ull;          // rounding offset for the top 32 bits
      sA[j * K + k] = turns_to_fix64(a0);
      sD[j * K + k] = turns_to_fix64((a1 - a0) / (double)hop);
      P += sf_frame_total(f_lo, f_hi, hop, inv_sr);
    }
  }
  __syncthreads();

  const float inv_hop = 1.0f / (float)hop;
  float* outb = out + (size_t)b * N + (size_t)i0 * hop;
  const int n_tile = nfr * hop;
  for (int lt = tid; lt < n_tile; lt += kSfThreads) {
    const int li = lt / hop, r = lt - li * hop;
    const float frac = (float)r * inv_hop;
    const float w1 = WINDOW ? (0.5f - 0.5f * cospif(frac)) : frac;
    const float w0 = 1.0f - w1;
    const unsigned long long c1 = (unsigned long long)(r + 1);
    const unsigned long long c2 = (unsigned long long)(((long long)r * (r + 1)) >> 1);
    const unsigned long long* P = sP + li * K;
    const unsigned long long* A = sA + li * K;
    const unsigned long long* D = sD + li * K;
    const float* f0r = sF + li * K;
    const float* f1r = f0r + K;
    const float* a0r = sAm + li * K;
    const float* a1r = a0r + K;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const unsigned long long ph = P[k] + c1 * A[k] + c2 * D[k];
      const float s = sinpif((float)(int)(uint32_t)(ph >> 32) * 4.656612873077393e-10f);
      const float lo = f0r[k], hi = f1r[k];
      const float fe = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), frac));   // core.py:617-620
      float amp = fmaf(a1r[k], w1, a0r[k] * w0);
      if (fe >= nyquist) amp = 0.f;                                         // core.py:888-890
      acc = fmaf(amp, s, acc);
    }
    if (accumulate) acc += outb[lt];
    outb[lt] = acc;
  }
}

}  // namespace ddsp
