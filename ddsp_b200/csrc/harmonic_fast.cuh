// Fast path of the fused harmonic kernel for hop % 64 == 0 (every reference
// config: hop = 64).  Same maths as harmonic.cuh; what changes is the mapping:
//
//   * one warp owns one frame at a time; a lane owns samples (r, r + 32) of a
//     64-sample chunk, so the two frame rows x0 = hd[i], x1 = hd[i+1] are
//     warp-uniform and come from shared memory as broadcast 128-bit loads;
//   * the frame slab hd[i0 .. i0+FT] is staged with ONE 1-D TMA bulk copy
//     (cp.async.bulk + mbarrier) when the rows are 16-byte multiples;
//   * sin(k phi) for k = 1..K comes from two interleaved Reinsch recurrences
//     (odd / even harmonics, angle 2 phi reduced to [-pi/2, pi/2]) held as one
//     packed f32x2 chain, so a harmonic pair costs 2 FFMA2 + 1 FFMA2 + 1 FADD2
//     per sample; amplitudes are NOT interpolated per oscillator - the two
//     frame rows get their own accumulators and the Hann / linear weights are
//     applied once per sample:  sum_k (w0 x0_k + w1 x1_k) s_k
//                             = w0 sum_k x0_k s_k + w1 sum_k x1_k s_k;
//   * the audio-rate Nyquist mask (core.py:942) is a per-sample live count;
//     harmonics below the warp-wide minimum run unmasked, the few between
//     min and max run with a per-lane predicate.
#pragma once
#include <cmath>
#include <cstdlib>

#include "harmonic.cuh"

namespace ddsp {

constexpr int kFastThreads = 256;
constexpr int kSinTabBits = 8;
constexpr int kSinTab = 1 << kSinTabBits;  // 256-entry (sin, cos) table

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  return __ffma2_rn(a, b, c);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  return __fadd2_rn(a, b);
}

struct FastSmem {
  size_t off_P, off_A, off_D, off_red, off_mbar, off_f0, off_amp, off_kc,
      off_w, off_tab, off_x, total;
};

__host__ __device__ inline FastSmem fast_smem_layout(int FT, int Kp, int hop) {
  FastSmem s;
  size_t o = 0;
  s.off_P = o;    o += sizeof(unsigned long long) * FT;
  s.off_A = o;    o += sizeof(unsigned long long) * FT;
  s.off_D = o;    o += sizeof(unsigned long long) * FT;
  s.off_red = o;  o += sizeof(unsigned long long) * 8;
  o = (o + 15) & ~(size_t)15;
  s.off_mbar = o; o += 16;
  s.off_tab = o;  o += sizeof(float2) * kSinTab;
  s.off_x = o;    o += sizeof(float) * (size_t)(FT + 1) * Kp;   // 16 B aligned
  s.off_f0 = o;   o += sizeof(float) * (FT + 1);
  s.off_amp = o;  o += sizeof(float) * (FT + 1);
  s.off_kc = o;   o += sizeof(int) * 2 * FT;
  s.off_w = o;    o += sizeof(float) * hop;
  s.total = (o + 15) & ~(size_t)15;
  return s;
}

// Slow, exact per-oscillator evaluation of one sample (frames with f0 < 1 Hz,
// where the live-count shortcut is not valid).
__device__ __noinline__ float harmonic_sample_exact(const float* x0,
                                                    const float* x1, float w0,
                                                    float w1, uint32_t p32,
                                                    float f_lo, float f_hi,
                                                    float frac, int K,
                                                    float nyq) {
  float acc = 0.f;
  uint32_t pk = 0;
  for (int k = 1; k <= K; ++k) {
    pk += p32;
    if (!(ref_harmonic_freq(f_lo, f_hi, frac, k) < nyq)) continue;
    float a = x0[k - 1] * w0 + x1[k - 1] * w1;
    acc = fmaf(a, sinpif((float)(int)pk * 4.656612873077393e-10f), acc);
  }
  return acc;
}

// Oscillator state for TWO samples (a, b) of the same frame, packed as f32x2
// (x = sample a, y = sample b).  Two Reinsch chains run in lock-step: chain 1
// holds the odd harmonics sin((1+2j) phi), chain 2 the even ones
// sin((2+2j) phi); both have angle 2 phi, reduced to [-pi/2, pi/2] - a half-turn
// shift flips the sign of every other step, which is why the accumulators are
// split by step parity (e = even j, o = odd j) and recombined with sigma.
struct Osc2 {
  float2 v1, v2, d1, d2;   // chain values and differences
  float2 nalpha;           // -4 sin^2(Phi_eff / 2) per sample
  float2 sigma;            // +1, or -1 where the chain angle was shifted
  float2 a0e1, a0e2, a0o1, a0o2;   // row x0: step parity x chain
  float2 a1e1, a1e2, a1o1, a1o2;   // row x1
};

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  return __fmul2_rn(a, b);
}
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }

__device__ __forceinline__ void osc2_init(Osc2& st, uint32_t pa, uint32_t pb,
                                          const float2* __restrict__ tab) {
  // sin/cos of the fundamental phase: table (top bits) + 3rd-order correction
  const uint32_t ia = (pa + (1u << (31 - kSinTabBits))) >> (32 - kSinTabBits);
  const uint32_t ib = (pb + (1u << (31 - kSinTabBits))) >> (32 - kSinTabBits);
  const int ra = (int)(pa - (ia << (32 - kSinTabBits)));
  const int rb = (int)(pb - (ib << (32 - kSinTabBits)));
  const float2 ta = tab[ia & (kSinTab - 1)];
  const float2 tb = tab[ib & (kSinTab - 1)];
  const float2 eps = fmul2(make_float2((float)ra, (float)rb),
                           bc2(1.4629180792671596e-9f));        // 2 pi / 2^32
  const float2 S = make_float2(ta.x, tb.x), C = make_float2(ta.y, tb.y);
  const float2 e2 = fmul2(eps, eps);
  // sin(t+e) = S (1 - e^2/2) + C e (1 - e^2/6); cos(t+e) = C (1 - e^2/2) - S e (1 - e^2/6)
  const float2 ce = ffma2(e2, bc2(-0.5f), bc2(1.0f));
  const float2 se = fmul2(eps, ffma2(e2, bc2(-0.16666667f), bc2(1.0f)));
  const float2 s1 = ffma2(C, se, fmul2(S, ce));
  const float2 c1 = ffma2(make_float2(-S.x, -S.y), se, fmul2(C, ce));
  const float2 ss = fmul2(s1, s1), cc = fmul2(c1, c1);
  const bool fa = ss.x > cc.x, fb = ss.y > cc.y;               // cos(2 phi) < 0
  const float2 s2 = fmul2(fadd2(s1, s1), c1);                  // sin(2 phi)
  st.nalpha = fmul2(make_float2(fminf(ss.x, cc.x), fminf(ss.y, cc.y)), bc2(-4.0f));
  st.v1 = s1;
  st.v2 = s2;
  st.d1 = make_float2(fa ? 0.0f : 2.0f * s1.x, fb ? 0.0f : 2.0f * s1.y);
  st.d2 = s2;
  st.sigma = make_float2(fa ? -1.0f : 1.0f, fb ? -1.0f : 1.0f);
  st.a0e1 = st.a0e2 = st.a0o1 = st.a0o2 = make_float2(0.f, 0.f);
  st.a1e1 = st.a1e2 = st.a1o1 = st.a1o2 = make_float2(0.f, 0.f);
}

// Advance both chains by one step (harmonics k, k+1 -> k+2, k+3).
__device__ __forceinline__ void osc2_step(Osc2& st) {
  st.d1 = ffma2(st.nalpha, st.v1, st.d1);
  st.d2 = ffma2(st.nalpha, st.v2, st.d2);
  st.v1 = fadd2(st.v1, st.d1);
  st.v2 = fadd2(st.v2, st.d2);
}

// One frame (hop samples, 64 per pass) of one batch item, executed by one warp.
// x0 / x1: the frame's two harmonic-distribution rows in shared memory.
__device__ __forceinline__ void harmonic_frame_pass(
    const float* __restrict__ x0, const float* __restrict__ x1,
    unsigned long long Pi, unsigned long long Ai, unsigned long long Di,
    int kc_a, int kc_b, float f_lo, float f_hi, float amp0, float amp1,
    const float* __restrict__ sW, const float2* __restrict__ sTab, int K,
    float nyquist, int hop, int lane, float* __restrict__ out, int accumulate) {
  const float inv_hop = 1.0f / (float)hop;
  for (int r0 = 0; r0 < hop; r0 += 64) {
    const int ra = r0 + lane, rb = ra + 32;
    // fundamental phase, 64-bit fixed point turns (inclusive cumsum)
    const unsigned long long pha = Pi + (unsigned long long)(ra + 1) * Ai +
        (unsigned long long)(((long long)ra * (ra + 1)) >> 1) * Di;
    const unsigned long long phb = Pi + (unsigned long long)(rb + 1) * Ai +
        (unsigned long long)(((long long)rb * (rb + 1)) >> 1) * Di;
    const uint32_t pa = (uint32_t)((pha + 0x80000000ull) >> 32);
    const uint32_t pb = (uint32_t)((phb + 0x80000000ull) >> 32);
    const float w1a_ = sW[ra], w1b_ = sW[rb];
    const float w0a = (1.0f - w1a_) * amp0, w1a = w1a_ * amp1;
    const float w0b = (1.0f - w1b_) * amp0, w1b = w1b_ * amp1;
    float ya, yb;
    if (kc_a < 0) {
      ya = harmonic_sample_exact(x0, x1, w0a, w1a, pa, f_lo, f_hi,
                                 (float)ra * inv_hop, K, nyquist);
      yb = harmonic_sample_exact(x0, x1, w0b, w1b, pb, f_lo, f_hi,
                                 (float)rb * inv_hop, K, nyquist);
    } else {
      int ka = kc_a, kb = kc_a, kmin = kc_a, kmax = kc_a;
      if (kc_a != kc_b) {      // live count changes inside this frame
        ka = live_harmonics(f_lo, f_hi, (float)ra * inv_hop, K, nyquist);
        kb = live_harmonics(f_lo, f_hi, (float)rb * inv_hop, K, nyquist);
        kmin = __reduce_min_sync(0xffffffffu, min(ka, kb));
        kmax = __reduce_max_sync(0xffffffffu, max(ka, kb));
      }
      Osc2 st;
      osc2_init(st, pa, pb, sTab);
      const int k_main = kmin & ~3;            // harmonics 1..k_main unmasked
      int k = 0;
#pragma unroll 2
      for (; k < k_main; k += 4) {
        const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
        st.a0e1 = ffma2(bc2(X0.x), st.v1, st.a0e1);
        st.a0e2 = ffma2(bc2(X0.y), st.v2, st.a0e2);
        st.a1e1 = ffma2(bc2(X1.x), st.v1, st.a1e1);
        st.a1e2 = ffma2(bc2(X1.y), st.v2, st.a1e2);
        osc2_step(st);
        st.a0o1 = ffma2(bc2(X0.z), st.v1, st.a0o1);
        st.a0o2 = ffma2(bc2(X0.w), st.v2, st.a0o2);
        st.a1o1 = ffma2(bc2(X1.z), st.v1, st.a1o1);
        st.a1o2 = ffma2(bc2(X1.w), st.v2, st.a1o2);
        osc2_step(st);
      }
      for (; k < kmax; k += 4) {               // masked tail (<= 2 passes)
        const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
        // harmonic numbers k+1 .. k+4; live iff number <= ka (sample a) / kb
        const float2 m1 = make_float2(k + 1 <= ka ? 1.f : 0.f, k + 1 <= kb ? 1.f : 0.f);
        const float2 m2 = make_float2(k + 2 <= ka ? 1.f : 0.f, k + 2 <= kb ? 1.f : 0.f);
        const float2 m3 = make_float2(k + 3 <= ka ? 1.f : 0.f, k + 3 <= kb ? 1.f : 0.f);
        const float2 m4 = make_float2(k + 4 <= ka ? 1.f : 0.f, k + 4 <= kb ? 1.f : 0.f);
        float2 u1 = fmul2(st.v1, m1), u2 = fmul2(st.v2, m2);
        st.a0e1 = ffma2(bc2(X0.x), u1, st.a0e1);
        st.a0e2 = ffma2(bc2(X0.y), u2, st.a0e2);
        st.a1e1 = ffma2(bc2(X1.x), u1, st.a1e1);
        st.a1e2 = ffma2(bc2(X1.y), u2, st.a1e2);
        osc2_step(st);
        u1 = fmul2(st.v1, m3); u2 = fmul2(st.v2, m4);
        st.a0o1 = ffma2(bc2(X0.z), u1, st.a0o1);
        st.a0o2 = ffma2(bc2(X0.w), u2, st.a0o2);
        st.a1o1 = ffma2(bc2(X1.z), u1, st.a1o1);
        st.a1o2 = ffma2(bc2(X1.w), u2, st.a1o2);
        osc2_step(st);
      }
      {
        const float2 r0s = ffma2(st.sigma, fadd2(st.a0o1, st.a0o2),
                                 fadd2(st.a0e1, st.a0e2));
        const float2 r1s = ffma2(st.sigma, fadd2(st.a1o1, st.a1o2),
                                 fadd2(st.a1e1, st.a1e2));
        const float2 y = ffma2(r1s, make_float2(w1a, w1b),
                               fmul2(r0s, make_float2(w0a, w0b)));
        ya = y.x;
        yb = y.y;
      }
    }
    float* o = out + r0;
    if (accumulate) {
      ya += o[lane];
      yb += o[lane + 32];
    }
    o[lane] = ya;
    o[lane + 32] = yb;
  }
}

// Harmonic.get_controls for up to four frame rows held in shared memory, by one
// warp (8 lanes per row): exp_sigmoid, frame-rate Nyquist mask on float32 f0*k,
// row normalisation with safe_divide (synths.py:110-117, core.py:894-907).  A
// row's live harmonics are a prefix (f0*k is monotone in k), so only
// ceil(live/4) float4 groups pay for exp_sigmoid - about a quarter of the row at
// the benchmark's f0 range.
__device__ __forceinline__ void harmonic_controls_rows(
    float* __restrict__ sX, const float* __restrict__ sF0, int r0, int rows_in,
    int K, int Kp, float nyquist, bool raw_scale, bool nyq, int lane) {
  const int K4 = Kp >> 2;                      // float4 groups per row
  const int sub = lane >> 3, l8 = lane & 7;
  const int r = r0 + sub;
  const bool row_ok = r < rows_in;
  float4* row4 = reinterpret_cast<float4*>(sX + (row_ok ? r : r0) * Kp);
  const float f = sF0[row_ok ? r : r0];
  int live = K;                                // harmonics with f0*k < sr/2
  if (nyq && f > 0.f) {
    int k = (int)fminf(nyquist / f, (float)K);
    while (k < K && __fmul_rn(f, (float)(k + 1)) < nyquist) ++k;
    while (k > 0 && !(__fmul_rn(f, (float)k) < nyquist)) --k;
    live = k;
  }
  float sum = 0.f;
  if (row_ok) {
    for (int c4 = l8; c4 < K4; c4 += 8) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * c4 < live) {
        v = row4[c4];
        float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float w = e[u];
          if (raw_scale) w = exp_sigmoid_f(w);
          if (4 * c4 + u >= live) w = 0.f;
          e[u] = w;
          sum += w;
        }
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
      row4[c4] = v;
    }
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.0f / ((sum == 0.0f) ? 1e-7f : sum);
  if (row_ok) {
    for (int c4 = l8; 4 * c4 < live; c4 += 8) {
      float4 v = row4[c4];
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
      row4[c4] = v;
    }
  }
}

template <bool WINDOW, int NT>
__global__ void __launch_bounds__(NT)
harmonic_fast_kernel(HarmonicParams p, int use_tma) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int FT = p.FT, Kp = p.Kp, K = p.K, F = p.F, hop = p.hop;
  const FastSmem L = fast_smem_layout(FT, Kp, hop);
  unsigned long long* sP = (unsigned long long*)(smem_raw + L.off_P);
  unsigned long long* sA = (unsigned long long*)(smem_raw + L.off_A);
  unsigned long long* sD = (unsigned long long*)(smem_raw + L.off_D);
  double* sRedD = (double*)(smem_raw + L.off_red);
  void* mbar = (void*)(smem_raw + L.off_mbar);
  float2* sTab = (float2*)(smem_raw + L.off_tab);
  float* sX = (float*)(smem_raw + L.off_x);
  float* sF0 = (float*)(smem_raw + L.off_f0);
  float* sAmp = (float*)(smem_raw + L.off_amp);
  int* sKc = (int*)(smem_raw + L.off_kc);
  float* sW = (float*)(smem_raw + L.off_w);

  const int b = blockIdx.y;
  const int i0 = blockIdx.x * FT;
  const int nfr = min(FT, F - i0);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const float* f0b = p.f0 + (size_t)b * F;
  const float* ampb = p.amps + (size_t)b * F;
  const int rows_in = min(nfr + 1, F - i0);

  // ---- 0. kick off the slab copy ----
  if (use_tma) {
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)rows_in * (uint32_t)K * 4u;
      mbar_expect_tx(mbar, bytes);
      tma_bulk_g2s(sX, p.hd + ((size_t)b * F + i0) * K, bytes, mbar);
    }
  }

  // ---- 1. phase at the start of this tile ----
  // sum_{j<i0} [hop a_j + (a_{j+1}-a_j)(hop-1)/2] telescopes to
  //   hop * sum_{j<i0} a_j + (hop-1)/2 * (a_{i0} - a_0),   a = f0 / sr,
  // so the prefix is one double-precision sum of f0 (<= 2^15 turns: 2^-38 turn
  // resolution), reduced over the CTA; every CTA recomputes it from f0 (<= 4 KB
  // of L2-resident data) instead of a serial scan over time.
  double part = 0.0;
  for (int j = tid; j < i0; j += NT) part += (double)f0b[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) sRedD[warp] = part;

  // ---- 2. small tables ----
  const bool raw_scale = p.ctl_flags & DDSP_B200_CTL_SCALE;
  for (int j = tid; j <= nfr; j += NT) {
    int g = min(i0 + j, F - 1);
    sF0[j] = f0b[g];
    const float a = ampb[g];
    sAmp[j] = raw_scale ? exp_sigmoid_f(a) : a;     // synths.py:110-111
  }
  for (int j = tid; j < kSinTab; j += NT) {
    float s, c;
    sincospif(2.0f * (float)j / (float)kSinTab, &s, &c);
    sTab[j] = make_float2(s, c);
  }
  {
    const float inv_hop = 1.0f / (float)hop;
    for (int r = tid; r < hop; r += NT) {
      const float frac = (float)r * inv_hop;
      sW[r] = WINDOW ? (0.5f - 0.5f * cospif(frac)) : frac;
    }
  }
  if (!use_tma) {
    if (p.hd != nullptr) {
      const float* hdb = p.hd + ((size_t)b * F + i0) * K;
      for (int idx = tid; idx < rows_in * Kp; idx += NT) {
        int r = idx / Kp, c = idx - r * Kp;
        sX[idx] = (c < K) ? hdb[r * K + c] : 0.f;
      }
    } else {
      for (int idx = tid; idx < rows_in * Kp; idx += NT)
        sX[idx] = (idx % Kp == 0) ? 1.0f : 0.f;
    }
  }
  __syncthreads();

  // ---- 3. per-frame phase tables + live counts at the frame ends ----
  // Frame totals are scanned with wrapping 64-bit adds (exact, associative), 32
  // frames per warp pass; P_i = tile prefix + exclusive scan of the totals.
  if (warp == 0) {
    double fsum = 0.0;
    for (int w = 0; w < NT / 32; ++w) fsum += sRedD[w];
    const double a_first = (double)f0b[0] * p.inv_sr;
    const double a_tile = (double)sF0[0] * p.inv_sr;
    unsigned long long P = turns_to_fix64(
        (double)hop * (fsum * p.inv_sr) + 0.5 * (hop - 1) * (a_tile - a_first));
    for (int base = 0; base < nfr; base += 32) {
      const int j = base + lane;
      unsigned long long tot = 0;
      if (j < nfr) {
        const double a0 = (double)sF0[j] * p.inv_sr;
        const double a1 = (double)sF0[j + 1] * p.inv_sr;
        sA[j] = turns_to_fix64(a0);
        sD[j] = turns_to_fix64((a1 - a0) / (double)hop);
        tot = turns_to_fix64((double)hop * a0 + (a1 - a0) * (0.5 * (hop - 1)));
      }
      unsigned long long incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (j < nfr) sP[j] = P + (incl - tot);
      P += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  for (int j = tid; j < nfr; j += NT) {
    const float f_lo = sF0[j], f_hi = sF0[j + 1];
    if (f_lo >= 1.0f && f_hi >= 1.0f) {
      sKc[2 * j] = live_harmonics(f_lo, f_hi, 0.0f, K, p.nyquist);
      sKc[2 * j + 1] = live_harmonics(
          f_lo, f_hi, (float)(hop - 1) * (1.0f / (float)hop), K, p.nyquist);
    } else {
      sKc[2 * j] = sKc[2 * j + 1] = -1;   // exact slow path
    }
  }
  if (use_tma) mbar_wait(mbar, 0);
  __syncthreads();
  if (p.ctl_flags != 0 && p.hd != nullptr) {
    // Harmonic.get_controls fused into the staging pass (synths.py:110-117):
    // exp_sigmoid, frame-rate Nyquist mask on float32 f0*k, row normalisation
    // with safe_divide.  One warp per frame row, rows stay in shared memory.
    const bool nyq = p.ctl_flags & DDSP_B200_CTL_NYQUIST;
    for (int r0 = warp * 4; r0 < rows_in; r0 += (NT / 32) * 4)
      harmonic_controls_rows(sX, sF0, r0, rows_in, K, Kp, p.nyquist, raw_scale, nyq,
                             lane);
    __syncthreads();
  }
  if (rows_in < nfr + 1) {                    // frame F := frame F-1
    for (int c = tid; c < Kp; c += NT)
      sX[nfr * Kp + c] = sX[(nfr - 1) * Kp + c];
    __syncthreads();
  }

  // ---- 4. samples: warp w takes frames w, w+8, ...; 64 samples per pass ----
  float* outb = p.audio + (size_t)b * p.N + (size_t)i0 * hop;
  for (int li = warp; li < nfr; li += NT / 32) {
    harmonic_frame_pass(sX + li * Kp, sX + (li + 1) * Kp, sP[li], sA[li], sD[li],
                        sKc[2 * li], sKc[2 * li + 1], sF0[li], sF0[li + 1],
                        sAmp[li], sAmp[li + 1], sW, sTab, K, p.nyquist, hop, lane,
                        outb + (size_t)li * hop, p.accumulate);
  }
}

inline bool harmonic_fast_supported(const HarmonicParams& p) {
  return (p.hop % 64 == 0) && p.hop <= 8192 && p.K <= 1024;
}

// Returns 0 on success, negative on error, 1 if it declines (caller falls back).
inline int launch_harmonic_fast(HarmonicParams p, cudaStream_t st) {
  p.Kp = (p.K + 3) & ~3;
  // debug / tuning knobs (not part of the ABI): threads per CTA and frames per tile
  static const int env_nt = [] { const char* e = getenv("DDSP_B200_HARM_NT"); return e ? atoi(e) : 128; }();
  static const int env_ft = [] { const char* e = getenv("DDSP_B200_HARM_FT"); return e ? atoi(e) : 0; }();
  // 4-warp CTAs (16 frames) measured 4 % faster than 8-warp ones (32 frames):
  // more resident CTAs, cheaper CTA-wide barriers (profiles/README.md).
  const int NT = (env_nt == 256) ? 256 : 128;
  int FT = std::max(1, 2048 / p.hop);
  if (NT == 128) FT = std::max(1, FT / 2);
  if (env_ft > 0) FT = std::min(FT, env_ft);
  const long long want_ctas = 4ll * kNumSMs;
  int ft_fill = (int)std::max<long long>(1, ((long long)p.B * p.F + want_ctas - 1) / want_ctas);
  FT = std::min(FT, std::max(ft_fill, std::min(8, p.F)));
  FT = std::min(FT, p.F);
  {
    // Wave quantisation: with ~3 resident CTAs per SM a grid of a few waves can
    // leave a third of the chip idle in the last one.  Halve the tile (at most
    // twice) when that buys more than 8 % of wave efficiency.
    const double slots = (NT == 128 ? 6.0 : 3.0) * kNumSMs;
    auto eff = [&](int ft) {
      const double n = (double)p.B * ((p.F + ft - 1) / ft);
      return n / (std::ceil(n / slots) * slots);
    };
    for (int tries = 0; tries < 2 && FT >= 16; ++tries) {
      if (eff(FT / 2) > eff(FT) + 0.08) FT /= 2; else break;
    }
  }
  while (FT > 1 && fast_smem_layout(FT, p.Kp, p.hop).total > 100 * 1024) FT = (FT + 1) / 2;
  const size_t smem = fast_smem_layout(FT, p.Kp, p.hop).total;
  if (smem > 200 * 1024) return 1;
  p.FT = FT;
  // TMA bulk copy needs 16-byte aligned rows and base
  const int use_tma = (p.hd != nullptr) && (p.K % 4 == 0) &&
                      (((uintptr_t)p.hd & 15) == 0);
  dim3 grid((p.F + FT - 1) / FT, p.B);
  cudaError_t e;
  if (p.amp_method == DDSP_B200_AMP_WINDOW) {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(harmonic_fast_kernel<true, 256>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(harmonic_fast_kernel<true, 128>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return 1;
    }
    if (NT == 128) harmonic_fast_kernel<true, 128><<<grid, 128, smem, st>>>(p, use_tma);
    else harmonic_fast_kernel<true, 256><<<grid, 256, smem, st>>>(p, use_tma);
  } else {
    if (smem > 48 * 1024) {
      e = cudaFuncSetAttribute(harmonic_fast_kernel<false, 256>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(harmonic_fast_kernel<false, 128>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return 1;
    }
    if (NT == 128) harmonic_fast_kernel<false, 128><<<grid, 128, smem, st>>>(p, use_tma);
    else harmonic_fast_kernel<false, 256><<<grid, 256, smem, st>>>(p, use_tma);
  }
  DDSP_CHECK_LAUNCH("harmonic_forward(fast)");
  return 0;
}

}  // namespace ddsp
