// harmonic_v4: fourth generation of the fused harmonic kernel (hop % 64 == 0).
// The mathematics is that of the earlier generations (closed-form phase, Reinsch
// chains over the harmonics with angle 2 phi, per-row accumulators, live-count
// Nyquist culling, get_controls fused into the slab staging - DESIGN.md 3.1).
// ncu on the third generation at B = 256 (profiles/r02_ncu_summary.txt): 425 warp
// instructions per 64-sample frame, 106 of them the oscillator loop, which already
// runs at the FMA pipe's packed rate (2 cycles per FFMA2) - everything else was
// overhead at one issue slot each.  What changed:
//
//   * the f32x2 lanes hold the SAME chain of the lane's TWO samples (r, r + 32),
//     not the (odd, even) chains of one sample.  Harmonic amplitudes enter as
//     broadcast scalar operands (FFMA2 R, R.F32, RR, RR); the seeds (table
//     look-up, rotation, double-angle, Reinsch constants), the weights and the
//     final combination are packed over the two samples: half the instructions,
//     no cross-half sums, eight accumulator registers less;
//   * the per-sample phase is two DFMAs on the otherwise idle FP64 pipe:
//     y = P' + c1 A + c2 D with P' = P + 1.5 * 2^20, whose low mantissa word IS
//     the 32-bit fixed-point phase (the frame phases P stay exact 64-bit
//     fixed point; the in-frame offset is < 2^13 turns, so the rounding is
//     <= 2^-32 turn).  Was 4 IMAD + 3 on the FMA / ALU pipes, twice per frame;
//   * get_controls writes each row once: exp_sigmoid on the live prefix, zeros
//     above it, and the row's 1 / sum goes into the frame amplitude (amp / sum),
//     not into a second pass over the row;
//   * frame-rate zeros make the Nyquist mask free: when every sample of a frame
//     has live count kc, row x0 is zero above kc by construction and row x1 is
//     zero above its own count; if that is <= kc the frame runs ceil(kc / 4)
//     unmasked groups.  Only the other frames take a masked group.
#pragma once
#include "harmonic_common.cuh"

namespace ddsp {
namespace hv4 {

#ifndef DDSP_HV4_NW
#define DDSP_HV4_NW 1
#endif
constexpr int NW = DDSP_HV4_NW;  // warps per CTA
constexpr int NT = NW * 32;

#ifndef DDSP_HV4_PHASE_F64
#define DDSP_HV4_PHASE_F64 0
#endif
#ifndef DDSP_HV4_MIN_CTAS
#define DDSP_HV4_MIN_CTAS (24 / DDSP_HV4_NW)
#endif

// 2^32 * (phase + 2^-9 turn): the table index is the top byte of the ROUNDED-UP
// phase, the residual its low 24 bits minus 2^23.
constexpr double kMagic = 1572864.0 + 0.001953125;          // 1.5 * 2^20 + 2^-9

struct __align__(16) FrameRec {
  unsigned long long P, A;       // F64 phase: bit patterns of doubles
  unsigned long long D;
  int ng;                        // > 0: uniform frame, ng unmasked groups; 0: per-sample
                                 // live counts; < 0: exact path (f0 < 1 Hz)
  int rem;                       // uniform frame: harmonics of the masked last group
  float f_lo, f_hi, amp0, amp1;  // amp = amplitude / row sum
};
static_assert(sizeof(FrameRec) == 48, "FrameRec must be three 16-byte words");

struct Smem {
  size_t off_mbar, off_sin, off_cos, off_x, off_w, off_red, off_inv, off_rec, off_live, total;
};

__host__ __device__ inline Smem smem_layout(int FW, int Kp, int hop) {
  Smem s;
  const size_t FT = (size_t)FW * NW;
  size_t o = 0;
  s.off_mbar = o; o += 16;
  s.off_sin = o;  o += sizeof(float) * kSinTab;
  s.off_cos = o;  o += sizeof(float) * kSinTab;
  s.off_x = o;    o += sizeof(float) * (FT + 1) * Kp;                 // 16 B aligned
  s.off_w = o;    o += (hop == 64) ? 0 : sizeof(float) * hop;
  o = (o + 15) & ~(size_t)15;
  s.off_red = o;  o += 16 * NW;                                       // double + u64 per warp
  s.off_rec = o;  o += sizeof(FrameRec) * FT;
  s.off_inv = o;  o += sizeof(float) * (FT + 1);
  s.off_live = o; o += sizeof(int) * (FT + 1);
  s.total = (o + 15) & ~(size_t)15;
  return s;
}

using hcm::phase32;

__device__ __forceinline__ float2 bffma2(float x, float2 v, float2 acc) {
  return __ffma2_rn(make_float2(x, x), v, acc);
}
__device__ __forceinline__ float2 bfmul2(float x, float2 v) {
  return __fmul2_rn(make_float2(x, x), v);
}

// Harmonic.get_controls for up to four rows (r0 .. r0+3 of this warp's block) in
// shared memory, 8 lanes per row (synths.py:110-117, core.py:894-907): exp_sigmoid
// on the live prefix, zeros above it, ONE store per element; the row's 1 / sum
// (safe_divide: a zero sum counts as 1e-7) goes to sInv[r] and is folded into the
// frame amplitude by the caller.
__device__ __forceinline__ void controls_rows4(float* __restrict__ sXw,
                                               const int* __restrict__ sLive,
                                               float* __restrict__ sInv, int r0,
                                               int nrows, int Kp, bool raw_scale,
                                               int lane) {
  const int K4 = Kp >> 2;
  const int sub = lane >> 3, l8 = lane & 7;
  const int r = r0 + sub;
  const bool row_ok = r < nrows;
  float4* row4 = reinterpret_cast<float4*>(sXw + (row_ok ? r : r0) * Kp);
  const int live = row_ok ? sLive[r] : 0;
  const int live4 = (live + 3) >> 2;                 // float4 groups with a live element
  float sum = 0.f;
  int c4 = l8;
  for (; c4 < live4; c4 += 8) {
    float4 x = row4[c4];
    if (raw_scale) {
      x.x = exp_sigmoid_f(x.x); x.y = exp_sigmoid_f(x.y);
      x.z = exp_sigmoid_f(x.z); x.w = exp_sigmoid_f(x.w);
    }
    const int k = 4 * c4;
    if (k + 1 >= live) x.y = 0.f;
    if (k + 2 >= live) x.z = 0.f;
    if (k + 3 >= live) x.w = 0.f;
    sum += (x.x + x.y) + (x.z + x.w);
    row4[c4] = x;
  }
  if (row_ok) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (; c4 < K4; c4 += 8) row4[c4] = z;
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  if (row_ok && l8 == 0) sInv[r] = __fdividef(1.0f, (sum == 0.0f) ? 1e-7f : sum);
}

// The exact per-oscillator slow path (f0 < 1 Hz) behind one call.
__device__ __noinline__ float sample_exact(const float* x0, int row_stride, float w0,
                                           float w1, uint32_t p32, float f_lo, float f_hi,
                                           float frac, int K, float nyq) {
  return harmonic_sample_exact(x0, x0 + row_stride, w0, w1, p32, f_lo, f_hi, frac, K, nyq);
}

// Oscillator state of the lane's TWO samples (.x = sample r, .y = sample r + 32):
// vo = sin((1+2j) phi), ve = sin((2+2j) phi); both chains step by the angle 2 phi
// reduced to [-pi/2, pi/2] (sg = -1 where it was shifted by half a turn: every
// other step then flips sign, hence the accumulators split by step parity e / o).
struct Osc2 {
  float2 vo, ve, dlo, dle, na, sg;
  float2 s0e, s0o, s1e, s1o;     // row x0 / x1, step parity: odd-harmonic chain
  float2 t0e, t0o, t1e, t1o;     // ... even-harmonic chain
};

// q = 2^32 * (phase + 2^-9) mod 2^32 for both samples.
__device__ __forceinline__ void osc_seed2(Osc2& st, uint32_t qa, uint32_t qb,
                                          const float* __restrict__ sSin,
                                          const float* __restrict__ sCos) {
  const uint32_t ia = qa >> (32 - kSinTabBits), ib = qb >> (32 - kSinTabBits);
  const float2 tx = make_float2(sSin[ia], sSin[ib]);
  const float2 ty = make_float2(sCos[ia], sCos[ib]);
  constexpr uint32_t kLow = (1u << (32 - kSinTabBits)) - 1u;
  constexpr float kC = 1.4629180792671596e-9f;                    // 2 pi / 2^32
  constexpr float kOff = -(float)(1u << (31 - kSinTabBits)) * kC;
  const float2 eps = __ffma2_rn(make_float2((float)(qa & kLow), (float)(qb & kLow)),
                                make_float2(kC, kC), make_float2(kOff, kOff));
  const float2 e2 = __fmul2_rn(eps, eps);
  const float2 ce = __ffma2_rn(e2, make_float2(-0.5f, -0.5f), make_float2(1.f, 1.f));
  const float2 se = __fmul2_rn(eps, __ffma2_rn(e2, make_float2(-0.16666667f, -0.16666667f),
                                               make_float2(1.f, 1.f)));
  const float2 nse = make_float2(-se.x, -se.y);
  const float2 s1 = __ffma2_rn(ty, se, __fmul2_rn(tx, ce));
  const float2 c1 = __ffma2_rn(tx, nse, __fmul2_rn(ty, ce));
  const float2 ss = __fmul2_rn(s1, s1), cc = __fmul2_rn(c1, c1);
  const float2 s2 = __fmul2_rn(__fadd2_rn(s1, s1), c1);           // sin(2 phi)
  st.sg = make_float2(ss.x > cc.x ? -1.0f : 1.0f, ss.y > cc.y ? -1.0f : 1.0f);
  st.na = __fmul2_rn(make_float2(fminf(ss.x, cc.x), fminf(ss.y, cc.y)),
                     make_float2(-4.0f, -4.0f));
  st.vo = s1;
  st.ve = s2;
  st.dlo = __ffma2_rn(s1, st.sg, s1);                             // 2 s1, or 0 if shifted
  st.dle = s2;
}

__device__ __forceinline__ void osc_step(Osc2& st) {
  st.dlo = __ffma2_rn(st.na, st.vo, st.dlo);
  st.dle = __ffma2_rn(st.na, st.ve, st.dle);
  st.vo = __fadd2_rn(st.vo, st.dlo);
  st.ve = __fadd2_rn(st.ve, st.dle);
}

// Four harmonics (k+1 .. k+4) of both samples: two chain steps.  Each of the
// eight accumulators takes ONE FFMA2 per group (two back-to-back updates of one
// accumulator made ptxas rotate registers through MOVs at the loop edge).
__device__ __forceinline__ void osc_group(Osc2& st, const float4& X0, const float4& X1) {
  st.s0e = bffma2(X0.x, st.vo, st.s0e);
  st.s1e = bffma2(X1.x, st.vo, st.s1e);
  st.t0e = bffma2(X0.y, st.ve, st.t0e);
  st.t1e = bffma2(X1.y, st.ve, st.t1e);
  osc_step(st);
  st.s0o = bffma2(X0.z, st.vo, st.s0o);
  st.s1o = bffma2(X1.z, st.vo, st.s1o);
  st.t0o = bffma2(X0.w, st.ve, st.t0o);
  st.t1o = bffma2(X1.w, st.ve, st.t1o);
  osc_step(st);
}

// First four harmonics: the accumulators are written, not accumulated into.
__device__ __forceinline__ void osc_group_first(Osc2& st, const float4& X0,
                                                const float4& X1) {
  st.s0e = bfmul2(X0.x, st.vo);
  st.s1e = bfmul2(X1.x, st.vo);
  st.t0e = bfmul2(X0.y, st.ve);
  st.t1e = bfmul2(X1.y, st.ve);
  osc_step(st);
  st.s0o = bfmul2(X0.z, st.vo);
  st.s1o = bfmul2(X1.z, st.vo);
  st.t0o = bfmul2(X0.w, st.ve);
  st.t1o = bfmul2(X1.w, st.ve);
  osc_step(st);
}

// Four harmonics with per-sample live counts (ka, kb): the sines are masked.
__device__ __forceinline__ void osc_group_masked(Osc2& st, const float4& X0,
                                                 const float4& X1, int k, int ka,
                                                 int kb) {
  float2 mo = make_float2(k + 1 <= ka ? st.vo.x : 0.f, k + 1 <= kb ? st.vo.y : 0.f);
  float2 me = make_float2(k + 2 <= ka ? st.ve.x : 0.f, k + 2 <= kb ? st.ve.y : 0.f);
  st.s0e = bffma2(X0.x, mo, st.s0e);
  st.s1e = bffma2(X1.x, mo, st.s1e);
  st.t0e = bffma2(X0.y, me, st.t0e);
  st.t1e = bffma2(X1.y, me, st.t1e);
  osc_step(st);
  mo = make_float2(k + 3 <= ka ? st.vo.x : 0.f, k + 3 <= kb ? st.vo.y : 0.f);
  me = make_float2(k + 4 <= ka ? st.ve.x : 0.f, k + 4 <= kb ? st.ve.y : 0.f);
  st.s0o = bffma2(X0.z, mo, st.s0o);
  st.s1o = bffma2(X1.z, mo, st.s1o);
  st.t0o = bffma2(X0.w, me, st.t0o);
  st.t1o = bffma2(X1.w, me, st.t1o);
  osc_step(st);
}

__device__ __forceinline__ float4 mask4u(const float4& X, int rem) {
  return make_float4(X.x, rem > 1 ? X.y : 0.f, rem > 2 ? X.z : 0.f, 0.f);
}

struct LaneConst {
#if DDSP_HV4_PHASE_F64
  double c1a, c2a, c1b, c2b;     // r + 1, r (r + 1) / 2 for the lane's two samples
#else
  uint32_t c1a, c2a, c1b, c2b;
#endif
  float2 w1;                     // amplitude weight of row x1 (Hann or linear)
};

template <bool WINDOW>
__device__ __forceinline__ LaneConst lane_const(int r0, int lane, float inv_hop,
                                                const float* __restrict__ sW) {
  LaneConst c;
  const uint32_t ra = r0 + lane, rb = ra + 32;
#if DDSP_HV4_PHASE_F64
  c.c1a = (double)(ra + 1); c.c2a = (double)((ra * (ra + 1)) >> 1);
  c.c1b = (double)(rb + 1); c.c2b = (double)((rb * (rb + 1)) >> 1);
#else
  c.c1a = ra + 1; c.c2a = (ra * (ra + 1)) >> 1;
  c.c1b = rb + 1; c.c2b = (rb * (rb + 1)) >> 1;
#endif
  // keep the constants in registers: ptxas otherwise re-derives them (and their
  // integer feeds) in every frame
#if DDSP_HV4_PHASE_F64
  asm volatile("" : "+d"(c.c1a), "+d"(c.c2a), "+d"(c.c1b), "+d"(c.c2b));
#else
  asm volatile("" : "+r"(c.c1a), "+r"(c.c2a), "+r"(c.c1b), "+r"(c.c2b));
#endif
  if (sW != nullptr) {
    c.w1 = make_float2(sW[ra], sW[rb]);
  } else {
    const float fa = (float)ra * inv_hop, fb = (float)rb * inv_hop;
    c.w1 = make_float2(WINDOW ? (0.5f - 0.5f * cospif(fa)) : fa,
                       WINDOW ? (0.5f - 0.5f * cospif(fb)) : fb);
  }
  asm volatile("" : "+f"(c.w1.x), "+f"(c.w1.y));
  return c;
}

// 64 samples of one frame (samples r0 + lane and r0 + lane + 32) by one warp.
// x0 = sX + xoff is the frame's own row, x1 = x0 + Kp the next one; `out` points at
// the lane's first sample.
__device__ __forceinline__ void frame_chunk(
    const float* __restrict__ sX, int xoff, int Kp, const FrameRec* __restrict__ rec,
    const LaneConst& lc, int r0, float inv_hop, const float* __restrict__ sSin,
    const float* __restrict__ sCos, int K, float nyquist, int lane,
    float* __restrict__ out, int accumulate) {
  const float* __restrict__ x0 = sX + xoff;
  const float* __restrict__ x1 = x0 + Kp;
  const ulonglong2 PA = *reinterpret_cast<const ulonglong2*>(&rec->P);
  const uint4 Dk = *reinterpret_cast<const uint4*>(&rec->D);
  const float4 fa = *reinterpret_cast<const float4*>(&rec->f_lo);
  const unsigned long long D = ((unsigned long long)Dk.y << 32) | Dk.x;
  const int ng = (int)Dk.z, rem = (int)Dk.w;
#if DDSP_HV4_PHASE_F64
  const double Pd = __longlong_as_double((long long)PA.x);
  const double Ad = __longlong_as_double((long long)PA.y);
  const double Dd = __longlong_as_double((long long)D);
  const uint32_t qa = (uint32_t)__double2loint(fma(lc.c2a, Dd, fma(lc.c1a, Ad, Pd)));
  const uint32_t qb = (uint32_t)__double2loint(fma(lc.c2b, Dd, fma(lc.c1b, Ad, Pd)));
#else
  const uint32_t qa = phase32(PA.x, PA.y, D, lc.c1a, lc.c2a);
  const uint32_t qb = phase32(PA.x, PA.y, D, lc.c1b, lc.c2b);
#endif
  // (1 - w1) amp0, w1 amp1
  const float2 w0 = __ffma2_rn(make_float2(-lc.w1.x, -lc.w1.y), make_float2(fa.z, fa.z),
                               make_float2(fa.z, fa.z));
  const float2 w1 = bfmul2(fa.w, lc.w1);
  float2 y;
  if (ng < 0) {              // f0 < 1 Hz somewhere: exact per-oscillator path
    constexpr uint32_t kRound = 1u << (31 - kSinTabBits);
    int ra = r0 + lane, xo = xoff;
    asm volatile("" : "+r"(ra), "+r"(xo));      // nothing of this path runs ahead of the branch
    const float fra = (float)ra * inv_hop, frb = (float)(ra + 32) * inv_hop;
    y.x = sample_exact(sX + xo, Kp, w0.x, w1.x, qa - kRound, fa.x, fa.y, fra, K, nyquist);
    y.y = sample_exact(sX + xo, Kp, w0.y, w1.y, qb - kRound, fa.x, fa.y, frb, K, nyquist);
  } else {
    Osc2 st;
    osc_seed2(st, qa, qb, sSin, sCos);
    if (ng > 0) {
      // every sample of the frame has the same live count: ng unmasked groups,
      // then (rem != 0) one group masked with warp-uniform predicates.  (One
      // loop from zeroed accumulators: peeling the first group to write the
      // accumulators cost ptxas 28 MOVs per frame at the merge points.)
      st.s0e = st.s0o = st.s1e = st.s1o = make_float2(0.f, 0.f);
      st.t0e = st.t0o = st.t1e = st.t1o = make_float2(0.f, 0.f);
      const int k_main = ng << 2;                   // ng >= 1
      int k = 0;
#pragma unroll 1
      do {
        const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
        osc_group(st, X0, X1);
        k += 4;
      } while (k < k_main);
      if (rem != 0) {
        const float4 X0 = mask4u(*reinterpret_cast<const float4*>(x0 + k), rem);
        const float4 X1 = mask4u(*reinterpret_cast<const float4*>(x1 + k), rem);
        osc_group(st, X0, X1);
      }
    } else {                 // live count changes inside this frame (or is < 4)
      int ra = r0 + lane;
      asm volatile("" : "+r"(ra));
      const float fra = (float)ra * inv_hop, frb = (float)(ra + 32) * inv_hop;
      const int ka = live_harmonics(fa.x, fa.y, fra, K, nyquist);
      const int kb = live_harmonics(fa.x, fa.y, frb, K, nyquist);
      const int kmin = __reduce_min_sync(0xffffffffu, min(ka, kb));
      const int kmax = __reduce_max_sync(0xffffffffu, max(ka, kb));
      st.s0e = st.s0o = st.s1e = st.s1o = make_float2(0.f, 0.f);
      st.t0e = st.t0o = st.t1e = st.t1o = make_float2(0.f, 0.f);
      const int k_main = kmin & ~3;
      int k = 0;
      for (; k < k_main; k += 4) {
        const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
        osc_group(st, X0, X1);
      }
      for (; k < kmax; k += 4) {
        const float4 X0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 X1 = *reinterpret_cast<const float4*>(x1 + k);
        osc_group_masked(st, X0, X1, k, ka, kb);
      }
    }
    const float2 t0 = __ffma2_rn(st.sg, __fadd2_rn(st.s0o, st.t0o), __fadd2_rn(st.s0e, st.t0e));
    const float2 t1 = __ffma2_rn(st.sg, __fadd2_rn(st.s1o, st.t1o), __fadd2_rn(st.s1e, st.t1e));
    y = __ffma2_rn(t1, w1, __fmul2_rn(t0, w0));
  }
  if (accumulate) {
    y.x += out[0];
    y.y += out[32];
  }
  out[0] = y.x;
  out[32] = y.y;
}

template <bool WINDOW, int HOPT>
__global__ void __launch_bounds__(NT, DDSP_HV4_MIN_CTAS)
harmonic_v4_kernel(HarmonicParams p, int use_tma, int FW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hop = HOPT ? HOPT : p.hop;
  const int Kp = p.Kp, K = p.K, F = p.F;
  const int FT = FW * NW;
  const Smem L = smem_layout(FW, Kp, hop);
  void* mbar = (void*)(smem_raw + L.off_mbar);
  float* sSin = (float*)(smem_raw + L.off_sin);
  float* sCos = (float*)(smem_raw + L.off_cos);
  float* sX = (float*)(smem_raw + L.off_x);
  float* sW = (HOPT == 64) ? nullptr : (float*)(smem_raw + L.off_w);
  double* sRedD = (double*)(smem_raw + L.off_red);                        // [NW]
  unsigned long long* sWarpTot = (unsigned long long*)(smem_raw + L.off_red) + NW;
  float* sInv = (float*)(smem_raw + L.off_inv);                           // [FT + 1]
  FrameRec* sRec = (FrameRec*)(smem_raw + L.off_rec);                     // [FT]
  int* sLive = (int*)(smem_raw + L.off_live);                             // [FT + 1]

  const int b = blockIdx.y;
  const int i0 = blockIdx.x * FT;
  const int nfr = min(FT, F - i0);
  const int rows_in = min(nfr + 1, F - i0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* f0b = p.f0 + (size_t)b * F;
  const float* ampb = p.amps + (size_t)b * F;

  // Programmatic dependent launch: the noise kernel of the decoder may start
  // its prologue on SMs this grid has vacated.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ---- 0. the frame slab: one TMA bulk copy, issued before anything else ----
  if (use_tma && tid == 0) {
    mbar_init(mbar, 1);
    const uint32_t bytes = (uint32_t)rows_in * (uint32_t)K * 4u;
    mbar_expect_tx(mbar, bytes);
    tma_bulk_g2s(sX, p.hd + ((size_t)b * F + i0) * K, bytes, mbar);
  }

  // ---- 1. every global load of the prologue is issued here, back to back: the
  //         f0 values before the tile (prefix sum, 16-byte loads where the item's
  //         row is aligned), the table, the tile's frames, the two frequencies of
  //         the closed-form tile phase.  One memory round trip instead of up to ten
  //         dependent ones (the prologue was 8 % of the instructions and 23 % of
  //         the warp time in the first capture of this kernel).
  //         RECORD WARPS: the per-frame quantities are computed with lane = frame
  //         by the first ceil(FT / 32) warps for the whole tile (every warp doing
  //         it for its own FW frames ran the same 300 instructions NW times).
  const int n_rec = (FT + 31) >> 5;                     // record warps
  const bool rec_warp = warp < n_rec;
  const int fr = warp * 32 + lane;                      // tile frame of a record lane
  const int cnt = max(0, min(32, nfr - warp * 32));     // frames of this record warp
  const bool raw_scale = p.ctl_flags & DDSP_B200_CTL_SCALE;
  float f = 0.f, a = 0.f, f_tile = 0.f, f_first = 0.f;
  float f_x = 0.f, a_x = 0.f;      // lane 31 of a full record warp: the frame after its range
  if (rec_warp && cnt > 0) {
    if (lane <= cnt) {
      const int g = min(i0 + fr, F - 1);                // frame F := frame F-1
      f = f0b[g];
      a = ampb[g];
    }
    if (lane == 31 && cnt == 32) {
      const int g = min(i0 + fr + 1, F - 1);
      f_x = f0b[g];
      a_x = ampb[g];
    }
    f_tile = f0b[i0];
    f_first = f0b[0];
  }
  constexpr int TPT = kSinTab / NT;                     // table entries per thread
  static_assert(TPT * NT == kSinTab, "the table splits evenly over the CTA");
  float2 tab[TPT];
#pragma unroll
  for (int u = 0; u < TPT; ++u) tab[u] = hcm::g_sincos256[tid + u * NT];
  double part = 0.0;
  if ((reinterpret_cast<uintptr_t>(f0b) & 15) == 0) {
    const float4* f4 = reinterpret_cast<const float4*>(f0b);
    const int n4 = i0 >> 2;
    float4 v[4];
    int j = tid;
    for (; j + 3 * NT < n4; j += 4 * NT) {
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = f4[j + u * NT];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        part += ((double)v[u].x + (double)v[u].y) + ((double)v[u].z + (double)v[u].w);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      v[u] = (j + u * NT < n4) ? f4[j + u * NT] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float tail = (tid < (i0 & 3)) ? f0b[4 * n4 + tid] : 0.f;   // i0 % 4 frames
#pragma unroll
    for (int u = 0; u < 4; ++u)
      part += ((double)v[u].x + (double)v[u].y) + ((double)v[u].z + (double)v[u].w);
    part += (double)tail;
  } else {
#pragma unroll 4
    for (int j = tid; j < i0; j += NT) part += (double)f0b[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) sRedD[warp] = part;
#pragma unroll
  for (int u = 0; u < TPT; ++u) {
    sSin[tid + u * NT] = tab[u].x;
    sCos[tid + u * NT] = tab[u].y;
  }
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

  // ---- 2. frame records (record warps, lane = frame) ----
  const bool have_ctl = (p.ctl_flags != 0) && (p.hd != nullptr);
  // rows are zero above their frame-rate live count (get_controls did it here)
  const bool zero_ok = have_ctl && (p.ctl_flags & DDSP_B200_CTL_NYQUIST);
  unsigned long long excl = 0;            // wrapping sum of the record warp's earlier frame totals
  if (rec_warp && cnt > 0) {
    if (raw_scale && lane <= cnt) a = exp_sigmoid_f(a);   // synths.py:110-111
    const float f_next = __shfl_down_sync(0xffffffffu, f, 1);
    const float a_next = __shfl_down_sync(0xffffffffu, a, 1);
    // lane 31 of a full record warp needs frame 32 of the NEXT record warp's range
    float f_n = f_next, a_n = a_next;
    if (lane == 31 && cnt == 32) {
      f_n = f_x;
      a_n = raw_scale ? exp_sigmoid_f(a_x) : a_x;
    }
    unsigned long long tot = 0;
    double a0 = 0.0, dd = 0.0;
    if (lane < cnt) {
      a0 = (double)f * p.inv_sr;
      const double a1 = (double)f_n * p.inv_sr;
      dd = (a1 - a0) * (1.0 / (double)hop);
      tot = turns_to_fix64((double)hop * a0 + (a1 - a0) * (0.5 * (hop - 1)));
    }
    unsigned long long incl = tot;                     // wrapping adds: exact
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    excl = incl - tot;
    if (lane == 31) sWarpTot[warp] = incl;             // total of the record warp's frames
    // frame-rate live count of a row (f0 * k < sr/2 in float32, core.py:888)
    auto row_live = [&](float fq) {
      int live = K;
      if ((p.ctl_flags & DDSP_B200_CTL_NYQUIST) && fq > 0.f) {
        int k = (int)fminf(p.nyquist / fq, (float)K);
        while (k < K && __fmul_rn(fq, (float)(k + 1)) < p.nyquist) ++k;
        while (k > 0 && !(__fmul_rn(fq, (float)k) < p.nyquist)) --k;
        live = k;
      }
      return live;
    };
    const int live = row_live(f);
    int live_next = __shfl_down_sync(0xffffffffu, live, 1);
    if (lane == 31 && cnt == 32) live_next = row_live(f_n);
    int ng = -1, rem = 0;                              // exact slow path
    if (lane < cnt && f >= 1.0f && f_n >= 1.0f) {
      const int kca = live_harmonics(f, f_n, 0.0f, K, p.nyquist);
      const int kcb = live_harmonics(f, f_n, (float)(hop - 1) * inv_hop, K, p.nyquist);
      ng = 0;                                          // per-sample live counts
      if (kca == kcb) {
        if (zero_ok && live == kca && live_next <= kca) {
          ng = max(1, (kca + 3) >> 2);                 // zeros above kca in both rows
        } else {
          ng = kca >> 2;                               // (0: the general path)
          rem = kca & 3;
        }
      }
    }
    if (lane < cnt) {
      FrameRec r;
#if DDSP_HV4_PHASE_F64
      r.P = 0;
      r.A = (unsigned long long)__double_as_longlong(a0);
      r.D = (unsigned long long)__double_as_longlong(dd);
#else
      r.P = 0; r.A = turns_to_fix64(a0); r.D = turns_to_fix64(dd);
#endif
      r.ng = ng; r.rem = rem;
      r.f_lo = f; r.f_hi = f_n; r.amp0 = a; r.amp1 = a_n;
      sRec[fr] = r;
    }
    if (lane <= cnt) sLive[fr] = live;
    if (lane == 31 && cnt == 32) sLive[fr + 1] = live_next;
  }
  __syncthreads();            // tables, mbarrier init, partial sums, records, (LDG slab)

  // phase at the start of the tile (telescoped closed form, one double-precision
  // evaluation: <= 2^15 turns, 2^-38 turn resolution), then the record warp's offset
  if (rec_warp && cnt > 0) {
    double base_sum = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) base_sum += sRedD[w];
    const double a_tile = (double)f_tile * p.inv_sr;
    const double a_first = (double)f_first * p.inv_sr;
    unsigned long long P0 = turns_to_fix64(
        (double)hop * (base_sum * p.inv_sr) + 0.5 * (hop - 1) * (a_tile - a_first));
    for (int w = 0; w < warp; ++w) P0 += sWarpTot[w];
    if (lane < cnt) {
      const unsigned long long Pf = P0 + excl;         // exact frame phase, 2^64 = 1 turn
#if DDSP_HV4_PHASE_F64
      // top 53 bits as a double in [0, 1), plus the magic that makes the low
      // mantissa word of P' + offsets the rounded 32-bit phase (+ 2^-9 turn)
      const double Pd = (double)(Pf >> 11) * 1.1102230246251565e-16 + kMagic;
      sRec[fr].P = (unsigned long long)__double_as_longlong(Pd);
#else
      sRec[fr].P = Pf + 0x80000000ull + (1ull << (63 - kSinTabBits));
#endif
    }
  }
  if (use_tma) mbar_wait(mbar, 0);

  // ---- 3. get_controls on the warp's own rows (synths.py:110-117) ----
  const int w0f = warp * FW;
  const int nfw = max(0, min(FW, nfr - w0f));
  if (nfw > 0) {
    float* sXw = sX + (size_t)w0f * Kp;
    const bool last = (w0f + nfw == nfr);
    int nrows = nfw;
    if (last && rows_in > nfr) nrows = nfw + 1;         // the real row after the tile
    if (have_ctl) {
      if (Kp <= 128) {
        for (int r0 = 0; r0 < nrows; r0 += 4)
          controls_rows4(sXw, sLive + w0f, sInv + w0f, r0, nrows, Kp, raw_scale, lane);
      } else {
        for (int r0 = 0; r0 < nrows; r0 += 4)
          hcm::controls_rows(sXw, sLive + w0f, r0, nrows, Kp, raw_scale, lane);
        for (int r = lane; r < nrows; r += 32) sInv[w0f + r] = 1.0f;
      }
    } else {
      for (int r = lane; r < nrows; r += 32) sInv[w0f + r] = 1.0f;
    }
    if (last && rows_in < nfr + 1) {                    // frame F := frame F-1
      __syncwarp();
      for (int c = lane; c < Kp; c += 32) sXw[nfw * Kp + c] = sXw[(nfw - 1) * Kp + c];
      if (lane == 0) sInv[w0f + nfw] = sInv[w0f + nfw - 1];
    }
  }
  __syncthreads();   // the row after a warp's block (and its 1 / sum) is its neighbour's;
                     // the records' phases

  // ---- 4. samples ----
  if (nfw > 0) {
    FrameRec* rec = sRec + w0f;
    if (lane < nfw) {                                   // amp / row sum
      rec[lane].amp0 *= sInv[w0f + lane];
      rec[lane].amp1 *= sInv[w0f + lane + 1];
    }
    __syncwarp();
    float* o = p.audio + (size_t)b * p.N + (size_t)(i0 + w0f) * hop + lane;
    int xoff = w0f * Kp;
    if (HOPT == 64) {
      const LaneConst lc = lane_const<WINDOW>(0, lane, inv_hop, nullptr);
#pragma unroll 1
      for (int li = 0; li < nfw; ++li, ++rec, xoff += Kp, o += 64) {
        frame_chunk(sX, xoff, Kp, rec, lc, 0, inv_hop, sSin, sCos, K, p.nyquist, lane, o,
                    p.accumulate);
      }
    } else {
      for (int li = 0; li < nfw; ++li, ++rec, xoff += Kp) {
        for (int r0 = 0; r0 < hop; r0 += 64, o += 64) {
          const LaneConst lc = lane_const<WINDOW>(r0, lane, inv_hop, sW);
          frame_chunk(sX, xoff, Kp, rec, lc, r0, inv_hop, sSin, sCos, K, p.nyquist, lane, o,
                      p.accumulate);
        }
      }
    }
  }
}

template <bool WINDOW, int HOPT>
inline cudaError_t launch_one(const HarmonicParams& p, int use_tma, int FW, dim3 grid,
                              size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(harmonic_v4_kernel<WINDOW, HOPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  harmonic_v4_kernel<WINDOW, HOPT><<<grid, NT, smem, st>>>(p, use_tma, FW);
  return cudaSuccess;
}

}  // namespace hv4

// Returns 0 on success, negative on error, 1 if the tile cannot fit shared memory
// (the caller then takes the generic kernel).
inline int launch_harmonic_v4(HarmonicParams p, cudaStream_t st) {
  using namespace hv4;
  p.Kp = (p.K + 3) & ~3;
  static const int env_fw = [] { const char* e = getenv("DDSP_B200_HARM_FW"); return e ? atoi(e) : 0; }();
  // One warp per CTA and 11 frames per warp (12 rows = three get_controls passes):
  // measured 295 us per B=256 decoder step against 301 for four warps x 8 frames
  // and 306 for four warps x 16.  Small grids shrink the tile until every SM has one.
  int FW = (NW == 1) ? 11 : 8;
  const long long want_ctas = 8ll * kNumSMs * (4 / NW);        // 32 warps per SM
  while (FW > 4 && (long long)p.B * ((p.F + FW * NW - 1) / (FW * NW)) < want_ctas) FW = (FW + 1) >> 1;
  while (FW > 1 && (long long)p.B * ((p.F + FW * NW - 1) / (FW * NW)) < kNumSMs) FW = (FW + 1) >> 1;
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
    (void)cudaGetLastError();
    set_error("harmonic_forward(v4): %s", cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  DDSP_CHECK_LAUNCH("harmonic_forward(v4)");
  return 0;
}

}  // namespace ddsp
