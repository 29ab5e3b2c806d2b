// harmonic_v2: the fused harmonic kernel for hop % 64 == 0, second generation.
// Same maths and the same bit-level decisions as harmonic_fast.cuh (read that
// header for the derivations: closed-form fixed-point phase, Reinsch chains over
// the harmonics, per-row accumulators, live-count Nyquist culling).  What changed
// is where the instructions go - the first kernel spent 3/4 of them outside the
// oscillator loop (ncu, profiles/r01_ncu_summary_v7.txt):
//
//   * WARP-AUTONOMOUS TILES.  A warp owns FW consecutive frames and builds
//     everything they need itself, one lane per frame: the phase prefix
//     (double-precision sum of f0 over the frames before it), the fixed-point
//     per-frame tables P/A/D (warp scan), the live counts, and get_controls on
//     its own rows.  The CTA meets twice (tables + TMA landed; controls done)
//     instead of five times, and no warp waits for warp 0.
//   * PER-SAMPLE PACKING.  f32x2 lanes now hold the (odd, even) harmonic chains
//     of ONE sample, so the row values (x_k, x_k+1) arrive from the LDS.128 as
//     the packed operand directly - no broadcast moves; a lane still owns two
//     samples (r, r + 32) for instruction-level parallelism, 4 accumulators each.
//   * CHEAP FIXED COSTS.  Lane constants (r + 1, r(r+1)/2, Hann weights) are
//     hoisted; the 64-bit phase is 4 IMADs per sample; the sin/cos table comes
//     from a constant array; the masked tail runs once with selects on the row
//     values; the frame-rate live count is computed once per row, not per pass.
#pragma once
#include "harmonic_fast.cuh"

namespace ddsp {
namespace hv2 {

constexpr int NW = 4;            // warps per CTA
constexpr int NT = NW * 32;

__device__ const float2 g_sincos256[kSinTab] = {
#include "sincos_tab.inc"
};

struct Smem {
  size_t off_mbar, off_tab, off_x, off_w, off_warp, warp_stride, total;
  // per-warp block (relative offsets)
  size_t w_P, w_A, w_D, w_f0, w_amp, w_kc, w_live;
};

__host__ __device__ inline Smem smem_layout(int FW, int Kp, int hop) {
  Smem s;
  size_t o = 0;
  s.off_mbar = o; o += 16;
  s.off_tab = o;  o += sizeof(float2) * kSinTab;
  s.off_x = o;    o += sizeof(float) * (size_t)(FW * NW + 1) * Kp;   // 16 B aligned
  s.off_w = o;    o += (hop == 64) ? 0 : sizeof(float) * hop;
  o = (o + 15) & ~(size_t)15;
  s.off_warp = o;
  size_t w = 0;
  s.w_P = w;    w += 8 * FW;
  s.w_A = w;    w += 8 * FW;
  s.w_D = w;    w += 8 * FW;
  s.w_f0 = w;   w += 4 * (FW + 1);
  s.w_amp = w;  w += 4 * (FW + 1);
  s.w_kc = w;   w += 8 * FW;
  s.w_live = w; w += 4 * (FW + 1);
  s.warp_stride = (w + 15) & ~(size_t)15;
  s.total = s.off_warp + NW * s.warp_stride;
  return s;
}

// top 32 bits of P + c1 * A + c2 * D (mod 2^64); P already carries the +2^31
// rounding offset.  4 IMADs.
__device__ __forceinline__ uint32_t phase32(unsigned long long P, unsigned long long A,
                                            unsigned long long D, uint32_t c1,
                                            uint32_t c2) {
  unsigned long long acc = P + (unsigned long long)(uint32_t)A * c1;
  uint32_t hi = (uint32_t)(acc >> 32) + (uint32_t)(A >> 32) * c1;
  acc = (((unsigned long long)hi << 32) | (uint32_t)acc) +
        (unsigned long long)(uint32_t)D * c2;
  return (uint32_t)(acc >> 32) + (uint32_t)(D >> 32) * c2;
}

// Oscillator state of ONE sample: .x = odd-harmonic chain sin((1+2j) phi),
// .y = even-harmonic chain sin((2+2j) phi); both step by the angle 2 phi reduced
// to [-pi/2, pi/2] (sigma = -1 where it was shifted by half a turn: every other
// step then flips sign, hence the accumulators split by step parity e / o).
struct Osc {
  float2 v, d, na;
  float sigma;
  float2 a0e, a0o, a1e, a1o;    // row x0 / x1, step parity
};

__device__ __forceinline__ void osc_init(Osc& st, uint32_t p,
                                         const float2* __restrict__ tab) {
  const uint32_t i = (p + (1u << (31 - kSinTabBits))) >> (32 - kSinTabBits);
  const int r = (int)(p - (i << (32 - kSinTabBits)));
  const float2 t = tab[i & (kSinTab - 1)];
  const float eps = (float)r * 1.4629180792671596e-9f;           // 2 pi / 2^32
  const float e2 = eps * eps;
  const float ce = fmaf(e2, -0.5f, 1.0f);
  const float se = eps * fmaf(e2, -0.16666667f, 1.0f);
  const float s1 = fmaf(t.y, se, t.x * ce);
  const float c1 = fmaf(-t.x, se, t.y * ce);
  const float ss = s1 * s1, cc = c1 * c1;
  const bool flip = ss > cc;                                     // cos(2 phi) < 0
  const float s2 = (s1 + s1) * c1;                               // sin(2 phi)
  const float na = -4.0f * fminf(ss, cc);
  st.v = make_float2(s1, s2);
  st.d = make_float2(flip ? 0.0f : s1 + s1, s2);
  st.na = make_float2(na, na);
  st.sigma = flip ? -1.0f : 1.0f;
  st.a0e = st.a0o = st.a1e = st.a1o = make_float2(0.f, 0.f);
}

// Four harmonics (k+1 .. k+4) of one sample: two chain steps.
__device__ __forceinline__ void osc_group(Osc& st, const float4& X0, const float4& X1) {
  st.a0e = ffma2(make_float2(X0.x, X0.y), st.v, st.a0e);
  st.a1e = ffma2(make_float2(X1.x, X1.y), st.v, st.a1e);
  st.d = ffma2(st.na, st.v, st.d);
  st.v = fadd2(st.v, st.d);
  st.a0o = ffma2(make_float2(X0.z, X0.w), st.v, st.a0o);
  st.a1o = ffma2(make_float2(X1.z, X1.w), st.v, st.a1o);
  st.d = ffma2(st.na, st.v, st.d);
  st.v = fadd2(st.v, st.d);
}

__device__ __forceinline__ float4 mask4(const float4& X, int k, int ks) {
  // harmonic numbers k+1 .. k+4 live iff number <= ks
  return make_float4(k + 1 <= ks ? X.x : 0.f, k + 2 <= ks ? X.y : 0.f,
                     k + 3 <= ks ? X.z : 0.f, k + 4 <= ks ? X.w : 0.f);
}

__device__ __forceinline__ float osc_finish(const Osc& st, float w0, float w1) {
  const float r0 = fmaf(st.sigma, st.a0o.x + st.a0o.y, st.a0e.x + st.a0e.y);
  const float r1 = fmaf(st.sigma, st.a1o.x + st.a1o.y, st.a1e.x + st.a1e.y);
  return fmaf(r1, w1, r0 * w0);
}

struct LaneConst {
  uint32_t c1a, c2a, c1b, c2b;   // r + 1, r (r + 1) / 2 for the lane's two samples
  float w1a, w1b;                // amplitude weight of row x1 (Hann or linear)
  float fra, frb;                // r / hop
};

template <bool WINDOW>
__device__ __forceinline__ LaneConst lane_const(int r0, int lane, float inv_hop,
                                                const float* __restrict__ sW) {
  LaneConst c;
  const uint32_t ra = r0 + lane, rb = ra + 32;
  c.c1a = ra + 1; c.c2a = (ra * (ra + 1)) >> 1;
  c.c1b = rb + 1; c.c2b = (rb * (rb + 1)) >> 1;
  c.fra = (float)ra * inv_hop;
  c.frb = (float)rb * inv_hop;
  if (sW != nullptr) {
    c.w1a = sW[ra];
    c.w1b = sW[rb];
  } else {
    c.w1a = WINDOW ? (0.5f - 0.5f * cospif(c.fra)) : c.fra;
    c.w1b = WINDOW ? (0.5f - 0.5f * cospif(c.frb)) : c.frb;
  }
  return c;
}

// 64 samples of one frame (samples r0 + lane and r0 + lane + 32) by one warp.
__device__ __forceinline__ void frame_chunk(
    const float* __restrict__ x0, const float* __restrict__ x1,
    unsigned long long Pr, unsigned long long A, unsigned long long D, int kc_a,
    int kc_b, float f_lo, float f_hi, float amp0, float amp1, const LaneConst& lc,
    const float2* __restrict__ sTab, int K, float nyquist, int lane,
    float* __restrict__ out, int accumulate) {
  const uint32_t pa = phase32(Pr, A, D, lc.c1a, lc.c2a);
  const uint32_t pb = phase32(Pr, A, D, lc.c1b, lc.c2b);
  const float w0a = (1.0f - lc.w1a) * amp0, w1a = lc.w1a * amp1;
  const float w0b = (1.0f - lc.w1b) * amp0, w1b = lc.w1b * amp1;
  float ya, yb;
  if (kc_a < 0) {            // f0 < 1 Hz somewhere: exact per-oscillator path
    ya = harmonic_sample_exact(x0, x1, w0a, w1a, pa, f_lo, f_hi, lc.fra, K, nyquist);
    yb = harmonic_sample_exact(x0, x1, w0b, w1b, pb, f_lo, f_hi, lc.frb, K, nyquist);
  } else {
    int ka = kc_a, kb = kc_a, kmin = kc_a, kmax = kc_a;
    if (kc_a != kc_b) {      // live count changes inside this frame
      ka = live_harmonics(f_lo, f_hi, lc.fra, K, nyquist);
      kb = live_harmonics(f_lo, f_hi, lc.frb, K, nyquist);
      kmin = __reduce_min_sync(0xffffffffu, min(ka, kb));
      kmax = __reduce_max_sync(0xffffffffu, max(ka, kb));
    }
    Osc sa, sb;
    osc_init(sa, pa, sTab);
    osc_init(sb, pb, sTab);
    const int k_main = kmin & ~3;              // harmonics 1..k_main unmasked
    int k = 0;
#pragma unroll 2
    for (; k < k_main; k += 4) {
      const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
      const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
      osc_group(sa, X0, X1);
      osc_group(sb, X0, X1);
    }
    for (; k < kmax; k += 4) {                 // masked tail (usually one pass)
      const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
      const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
      osc_group(sa, mask4(X0, k, ka), mask4(X1, k, ka));
      osc_group(sb, mask4(X0, k, kb), mask4(X1, k, kb));
    }
    ya = osc_finish(sa, w0a, w1a);
    yb = osc_finish(sb, w0b, w1b);
  }
  if (accumulate) {
    ya += out[lane];
    yb += out[lane + 32];
  }
  out[lane] = ya;
  out[lane + 32] = yb;
}

// Harmonic.get_controls for up to four rows (r0 .. r0+3 of this warp's block) in
// shared memory, 8 lanes per row: exp_sigmoid on the live prefix, zeros above it,
// row normalisation with safe_divide (synths.py:110-117, core.py:894-907).  The
// frame-rate live count of each row (f0*k < sr/2 in float32) was computed once
// per row by the caller.
__device__ __forceinline__ void controls_rows(float* __restrict__ sXw,
                                              const int* __restrict__ sLive, int r0,
                                              int nrows, int Kp, bool raw_scale,
                                              int lane) {
  const int K4 = Kp >> 2;
  const int sub = lane >> 3, l8 = lane & 7;
  const int r = r0 + sub;
  const bool row_ok = r < nrows;
  float4* row4 = reinterpret_cast<float4*>(sXw + (row_ok ? r : r0) * Kp);
  const int live = sLive[row_ok ? r : r0];
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

#ifndef DDSP_HV2_MIN_CTAS
#define DDSP_HV2_MIN_CTAS 4
#endif
template <bool WINDOW, int HOPT>
__global__ void __launch_bounds__(NT, DDSP_HV2_MIN_CTAS)
harmonic_v2_kernel(HarmonicParams p, int use_tma, int FW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hop = HOPT ? HOPT : p.hop;
  const int Kp = p.Kp, K = p.K, F = p.F;
  const int FT = FW * NW;
  const Smem L = smem_layout(FW, Kp, hop);
  void* mbar = (void*)(smem_raw + L.off_mbar);
  float2* sTab = (float2*)(smem_raw + L.off_tab);
  float* sX = (float*)(smem_raw + L.off_x);
  float* sW = (HOPT == 64) ? nullptr : (float*)(smem_raw + L.off_w);

  const int b = blockIdx.y;
  const int i0 = blockIdx.x * FT;
  const int nfr = min(FT, F - i0);
  const int rows_in = min(nfr + 1, F - i0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* f0b = p.f0 + (size_t)b * F;
  const float* ampb = p.amps + (size_t)b * F;
  unsigned char* wbase = smem_raw + L.off_warp + warp * L.warp_stride;
  unsigned long long* sP = (unsigned long long*)(wbase + L.w_P);
  unsigned long long* sA = (unsigned long long*)(wbase + L.w_A);
  unsigned long long* sD = (unsigned long long*)(wbase + L.w_D);
  float* sF0 = (float*)(wbase + L.w_f0);
  float* sAmp = (float*)(wbase + L.w_amp);
  int2* sKc = (int2*)(wbase + L.w_kc);
  int* sLive = (int*)(wbase + L.w_live);

  // Programmatic dependent launch: the next kernel in the stream (the noise
  // kernel of the decoder) may start its prologue on SMs this grid has vacated.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ---- 0. the frame slab: one TMA bulk copy, issued before anything else ----
  if (use_tma && tid == 0) {
    mbar_init(mbar, 1);
    const uint32_t bytes = (uint32_t)rows_in * (uint32_t)K * 4u;
    mbar_expect_tx(mbar, bytes);
    tma_bulk_g2s(sX, p.hd + ((size_t)b * F + i0) * K, bytes, mbar);
  }

  // ---- 1. CTA-wide tables ----
  for (int j = tid; j < kSinTab; j += NT) sTab[j] = g_sincos256[j];
  const float inv_hop = 1.0f / (float)hop;
  if (HOPT != 64) {
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

  // ---- 2. this warp's frames: w0f .. w0f + nfw - 1 (lane = frame) ----
  const int w0f = warp * FW;
  const int nfw = max(0, min(FW, nfr - w0f));
  const bool raw_scale = p.ctl_flags & DDSP_B200_CTL_SCALE;
  const bool have_ctl = (p.ctl_flags != 0) && (p.hd != nullptr);
  if (nfw > 0) {
    const int g0 = i0 + w0f;
    // phase at the start of the warp's block: sum_{j<g0} [hop a_j + (a_{j+1}-a_j)
    // (hop-1)/2] telescopes to hop * sum a_j + (hop-1)/2 (a_g0 - a_0), a = f0/sr;
    // one double-precision sum of f0 (<= 2^15 turns: 2^-38 turn resolution).
    double part = 0.0;
    for (int j = lane; j < g0; j += 32) part += (double)f0b[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const int g = min(g0 + lane, F - 1);              // frame F := frame F-1
    float f = 0.f, a = 0.f;
    if (lane <= nfw) {
      f = f0b[g];
      a = ampb[g];
      if (raw_scale) a = exp_sigmoid_f(a);              // synths.py:110-111
    }
    const float f_next = __shfl_down_sync(0xffffffffu, f, 1);
    const float f_first = __shfl_sync(0xffffffffu, f, 0);
    const double a_tile = (double)f_first * p.inv_sr;
    const double a_first = (double)f0b[0] * p.inv_sr;
    const unsigned long long P0 = turns_to_fix64(
        (double)hop * (part * p.inv_sr) + 0.5 * (hop - 1) * (a_tile - a_first));
    unsigned long long tot = 0, Af = 0, Df = 0;
    if (lane < nfw) {
      const double a0 = (double)f * p.inv_sr;
      const double a1 = (double)f_next * p.inv_sr;
      Af = turns_to_fix64(a0);
      Df = turns_to_fix64((a1 - a0) / (double)hop);
      tot = turns_to_fix64((double)hop * a0 + (a1 - a0) * (0.5 * (hop - 1)));
    }
    unsigned long long incl = tot;                     // wrapping adds: exact
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    // frame-rate live count of the row (f0 * k < sr/2 in float32, core.py:888)
    int live = K;
    if ((p.ctl_flags & DDSP_B200_CTL_NYQUIST) && f > 0.f) {
      int k = (int)fminf(p.nyquist / f, (float)K);
      while (k < K && __fmul_rn(f, (float)(k + 1)) < p.nyquist) ++k;
      while (k > 0 && !(__fmul_rn(f, (float)k) < p.nyquist)) --k;
      live = k;
    }
    int kca = -1, kcb = -1;                            // exact slow path
    if (lane < nfw && f >= 1.0f && f_next >= 1.0f) {
      kca = live_harmonics(f, f_next, 0.0f, K, p.nyquist);
      kcb = live_harmonics(f, f_next, (float)(hop - 1) * inv_hop, K, p.nyquist);
    }
    if (lane < nfw) {
      sP[lane] = P0 + (incl - tot) + 0x80000000ull;    // rounding offset folded in
      sA[lane] = Af;
      sD[lane] = Df;
      sKc[lane] = make_int2(kca, kcb);
    }
    if (lane <= nfw) {
      sF0[lane] = f;
      sAmp[lane] = a;
      sLive[lane] = live;
    }
  }
  __syncthreads();            // tables, mbarrier init, (LDG-staged slab) visible
  if (use_tma) mbar_wait(mbar, 0);

  // ---- 3. get_controls on the warp's own rows (synths.py:110-117) ----
  const bool need_sync2 = have_ctl || (rows_in < nfr + 1);
  if (nfw > 0) {
    float* sXw = sX + (size_t)w0f * Kp;
    const bool last = (w0f + nfw == nfr);
    int nrows = nfw;
    if (last && rows_in > nfr) nrows = nfw + 1;         // the real row after the tile
    if (have_ctl) {
      for (int r0 = 0; r0 < nrows; r0 += 4)
        controls_rows(sXw, sLive, r0, nrows, Kp, raw_scale, lane);
    }
    if (last && rows_in < nfr + 1) {                    // frame F := frame F-1
      __syncwarp();
      for (int c = lane; c < Kp; c += 32) sXw[nfw * Kp + c] = sXw[(nfw - 1) * Kp + c];
    }
  }
  if (need_sync2) __syncthreads();   // the row after a warp's block is its neighbour's

  // ---- 4. samples ----
  if (nfw > 0) {
    float* outw = p.audio + (size_t)b * p.N + (size_t)(i0 + w0f) * hop;
    const float* xw = sX + (size_t)w0f * Kp;
    if (HOPT == 64) {
      const LaneConst lc = lane_const<WINDOW>(0, lane, inv_hop, nullptr);
      for (int li = 0; li < nfw; ++li) {
        const int2 kc = sKc[li];
        frame_chunk(xw + li * Kp, xw + (li + 1) * Kp, sP[li], sA[li], sD[li], kc.x,
                    kc.y, sF0[li], sF0[li + 1], sAmp[li], sAmp[li + 1], lc, sTab, K,
                    p.nyquist, lane, outw + (size_t)li * 64, p.accumulate);
      }
    } else {
      for (int li = 0; li < nfw; ++li) {
        const int2 kc = sKc[li];
        for (int r0 = 0; r0 < hop; r0 += 64) {
          const LaneConst lc = lane_const<WINDOW>(r0, lane, inv_hop, sW);
          frame_chunk(xw + li * Kp, xw + (li + 1) * Kp, sP[li], sA[li], sD[li], kc.x,
                      kc.y, sF0[li], sF0[li + 1], sAmp[li], sAmp[li + 1], lc, sTab, K,
                      p.nyquist, lane, outw + (size_t)li * hop + r0, p.accumulate);
        }
      }
    }
  }
}

template <bool WINDOW, int HOPT>
inline cudaError_t launch_one(const HarmonicParams& p, int use_tma, int FW, dim3 grid,
                              size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(harmonic_v2_kernel<WINDOW, HOPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  harmonic_v2_kernel<WINDOW, HOPT><<<grid, NT, smem, st>>>(p, use_tma, FW);
  return cudaSuccess;
}

}  // namespace hv2

// Returns 0 on success, negative on error, 1 if it declines (caller falls back).
inline int launch_harmonic_v2(HarmonicParams p, cudaStream_t st) {
  using namespace hv2;
  p.Kp = (p.K + 3) & ~3;
  // tuning knob (not part of the ABI): frames per warp
  static const int env_fw = [] { const char* e = getenv("DDSP_B200_HARM_FW"); return e ? atoi(e) : 0; }();
  // Frames per warp: 8 amortises the per-warp prologue best; fewer when the grid
  // would not fill the chip a few times over (the CTAs' work varies ~10x with f0,
  // so several waves are needed for the block scheduler to balance it).
  // (measured, B200, HBM-cold inputs: B=256 -> FW 16 / 8 / 4 / 2 = 163 / 163 / 183 /
  // 238 us; B=32 -> 37.8 / 34.9 / 31.1 / 35.0 us)
  int FW = 16;
  const long long want_ctas = 8ll * kNumSMs;
  while (FW > 4 && (long long)p.B * ((p.F + FW * NW - 1) / (FW * NW)) < want_ctas) FW >>= 1;
  while (FW > 1 && (long long)p.B * ((p.F + FW * NW - 1) / (FW * NW)) < kNumSMs) FW >>= 1;
  if (env_fw > 0) FW = std::min(32, env_fw);
  FW = std::max(1, std::min(FW, (p.F + NW - 1) / NW));
  while (FW > 1 && smem_layout(FW, p.Kp, p.hop).total > 64 * 1024) FW = (FW + 1) / 2;
  const size_t smem = smem_layout(FW, p.Kp, p.hop).total;
  if (smem > 200 * 1024) return 1;
  const int use_tma = (p.hd != nullptr) && (p.K % 4 == 0) &&
                      (((uintptr_t)p.hd & 15) == 0);
  dim3 grid((p.F + FW * NW - 1) / (FW * NW), p.B);
  cudaError_t e;
  const bool win = p.amp_method == DDSP_B200_AMP_WINDOW;
  if (p.hop == 64) {
    e = win ? launch_one<true, 64>(p, use_tma, FW, grid, smem, st)
            : launch_one<false, 64>(p, use_tma, FW, grid, smem, st);
  } else {
    e = win ? launch_one<true, 0>(p, use_tma, FW, grid, smem, st)
            : launch_one<false, 0>(p, use_tma, FW, grid, smem, st);
  }
  if (e != cudaSuccess) {
    // a real error (attribute / launch configuration): report it - falling back to
    // the generic kernel here would hide a 10x slowdown behind a correct result
    (void)cudaGetLastError();
    set_error("harmonic_forward(v2): %s", cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  DDSP_CHECK_LAUNCH("harmonic_forward(v2)");
  return 0;
}

}  // namespace ddsp
