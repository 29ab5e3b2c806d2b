// harmonic_v3: third generation of the fused harmonic kernel (hop % 64 == 0).
// Maths and every bit-level decision are those of the first two generations
// (profiles/experiments/harmonic_fast.cuh.txt, harmonic_v2.cuh.txt; shared pieces in
// harmonic_common.cuh: closed-form 64-bit fixed-point phase, Reinsch chains over the
// harmonics in (odd, even) f32x2 lanes, per-row accumulators, live-count Nyquist
// culling, get_controls fused into the slab staging).  What changed is the
// instruction budget OUTSIDE the oscillator loop - ncu on v2 at B = 256
// (profiles/r01_ncu_summary_v8.txt) showed 472 warp instructions per 64-sample
// frame of which only 168 were the oscillator loop:
//
//   * one 48-byte FRAME RECORD per frame (P, A | D, kc_a, kc_b | f_lo, f_hi,
//     amp0, amp1) read with three LDS.128 instead of seven scalar loads with
//     their own address arithmetic;
//   * the accumulators are initialised by the first harmonic group (FMUL2) -
//     v2 zeroed sixteen registers, twice (the compiler re-materialised them);
//   * when every sample of the frame has the same live count (9 frames in 10)
//     the last partial group is masked with WARP-UNIFORM predicates (12
//     instructions) instead of the per-sample masked tail (50);
//   * the phase prefix is summed once per CTA (128 threads, <= 8 loads each) and
//     completed per warp with a wrapping 64-bit scan, instead of every warp
//     re-summing f0[0 .. g0) from global memory in double precision;
//   * per-lane constants and the Hann weights come out of the frame loop.
#pragma once
#include "harmonic_common.cuh"

namespace ddsp {
namespace hv3 {

constexpr int NW = 4;            // warps per CTA
constexpr int NT = NW * 32;

struct __align__(16) FrameRec {
  unsigned long long P, A;       // P carries the +2^31 rounding offset
  unsigned long long D;
  int kca, kcb;                  // live counts at r = 0 / r = hop-1; kca < 0: exact path
  float f_lo, f_hi, amp0, amp1;
};
static_assert(sizeof(FrameRec) == 48, "FrameRec must be three 16-byte words");

struct Smem {
  size_t off_mbar, off_tab, off_x, off_w, off_red, off_warp, warp_stride, total;
  size_t w_rec, w_live;
};

__host__ __device__ inline Smem smem_layout(int FW, int Kp, int hop) {
  Smem s;
  size_t o = 0;
  s.off_mbar = o; o += 16;
  s.off_tab = o;  o += sizeof(float2) * kSinTab;
  s.off_x = o;    o += sizeof(float) * (size_t)(FW * NW + 1) * Kp;   // 16 B aligned
  s.off_w = o;    o += (hop == 64) ? 0 : sizeof(float) * hop;
  o = (o + 15) & ~(size_t)15;
  s.off_red = o;  o += 16 * NW;                                       // double + u64 per warp
  s.off_warp = o;
  size_t w = 0;
  s.w_rec = w;  w += sizeof(FrameRec) * FW;
  s.w_live = w; w += 4 * (FW + 1);
  s.warp_stride = (w + 15) & ~(size_t)15;
  s.total = s.off_warp + NW * s.warp_stride;
  return s;
}

using hcm::Osc;
using hcm::osc_group;
using hcm::osc_finish;
using hcm::phase32;
using hcm::mask4;

// Harmonic.get_controls for up to four rows (r0 .. r0+3 of this warp's block) in
// shared memory, 8 lanes per row: exp_sigmoid on the live prefix, zeros above it,
// row normalisation with safe_divide (synths.py:110-117, core.py:894-907).  Same
// arithmetic as hcm::controls_rows; the values stay in registers between the sum
// and the normalisation (one store instead of store / load / store), only as many
// 8-group passes run as the longest of the four rows needs (usually one), and the
// zero fill of the masked tail is a separate tight loop.  Rows of up to 128
// harmonics; wider rows take hcm::controls_rows.
__device__ __forceinline__ void controls_rows4(float* __restrict__ sXw,
                                               const int* __restrict__ sLive, int r0,
                                               int nrows, int Kp, bool raw_scale,
                                               int lane) {
  const int K4 = Kp >> 2;
  const int sub = lane >> 3, l8 = lane & 7;
  const int r = r0 + sub;
  const bool row_ok = r < nrows;
  float4* row4 = reinterpret_cast<float4*>(sXw + (row_ok ? r : r0) * Kp);
  const int live = row_ok ? sLive[r] : 0;
  const int live4 = (live + 3) >> 2;                 // float4 groups with a live element
  const int n_it = (__reduce_max_sync(0xffffffffu, live4) + 7) >> 3;   // warp-uniform
  float4 v[4];
  float sum = 0.f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (it < n_it) {
      const int c4 = l8 + 8 * it;
      if (c4 < live4) {
        const float4 x = row4[c4];
        float e[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float w = e[u];
          if (raw_scale) w = exp_sigmoid_f(w);
          if (4 * c4 + u >= live) w = 0.f;
          e[u] = w;
          sum += w;
        }
        v[it] = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 4);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.0f / ((sum == 0.0f) ? 1e-7f : sum);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if (it < n_it) {
      const int c4 = l8 + 8 * it;
      if (c4 < live4)
        row4[c4] = make_float4(v[it].x * inv, v[it].y * inv, v[it].z * inv, v[it].w * inv);
    }
  }
  if (row_ok) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c4 = live4 + l8; c4 < K4; c4 += 8) row4[c4] = z;
  }
}

// osc_init without zeroing the accumulators (the first group writes them).
__device__ __forceinline__ void osc_seed(Osc& st, uint32_t p,
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
}

// First four harmonics: the accumulators are written, not accumulated into.
__device__ __forceinline__ void osc_group_first(Osc& st, const float4& X0,
                                                const float4& X1) {
  st.a0e = __fmul2_rn(make_float2(X0.x, X0.y), st.v);
  st.a1e = __fmul2_rn(make_float2(X1.x, X1.y), st.v);
  st.d = ffma2(st.na, st.v, st.d);
  st.v = fadd2(st.v, st.d);
  st.a0o = __fmul2_rn(make_float2(X0.z, X0.w), st.v);
  st.a1o = __fmul2_rn(make_float2(X1.z, X1.w), st.v);
  st.d = ffma2(st.na, st.v, st.d);
  st.v = fadd2(st.v, st.d);
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
    const FrameRec* __restrict__ rec, const LaneConst& lc,
    const float2* __restrict__ sTab, int K, float nyquist, int lane,
    float* __restrict__ out, int accumulate) {
  const ulonglong2 PA = *reinterpret_cast<const ulonglong2*>(&rec->P);
  const uint4 Dk = *reinterpret_cast<const uint4*>(&rec->D);
  const float4 fa = *reinterpret_cast<const float4*>(&rec->f_lo);
  const unsigned long long D = ((unsigned long long)Dk.y << 32) | Dk.x;
  const int kc_a = (int)Dk.z, kc_b = (int)Dk.w;
  const uint32_t pa = phase32(PA.x, PA.y, D, lc.c1a, lc.c2a);
  const uint32_t pb = phase32(PA.x, PA.y, D, lc.c1b, lc.c2b);
  const float w1a = lc.w1a * fa.w, w0a = fmaf(-lc.w1a, fa.z, fa.z);   // (1 - w1) amp0
  const float w1b = lc.w1b * fa.w, w0b = fmaf(-lc.w1b, fa.z, fa.z);
  float ya, yb;
  if (kc_a < 0) {            // f0 < 1 Hz somewhere: exact per-oscillator path
    ya = harmonic_sample_exact(x0, x1, w0a, w1a, pa, fa.x, fa.y, lc.fra, K, nyquist);
    yb = harmonic_sample_exact(x0, x1, w0b, w1b, pb, fa.x, fa.y, lc.frb, K, nyquist);
  } else {
    Osc sa, sb;
    osc_seed(sa, pa, sTab);
    osc_seed(sb, pb, sTab);
    if (kc_a == kc_b) {
      // every sample of the frame has the same live count kc: ceil(kc / 4)
      // groups, the last one masked with warp-uniform predicates
      const int kc = kc_a;
      const int g_full = kc >> 2;                 // fully live groups
      int k = 0;
      if (g_full > 0) {
        const float4 X0 = *reinterpret_cast<const float4*>(x0);
        const float4 X1 = *reinterpret_cast<const float4*>(x1);
        osc_group_first(sa, X0, X1);
        osc_group_first(sb, X0, X1);
        k = 4;
        const int k_main = g_full << 2;
#pragma unroll 2
        for (; k < k_main; k += 4) {
          const float4 Y0 = *reinterpret_cast<const float4*>(x0 + k);
          const float4 Y1 = *reinterpret_cast<const float4*>(x1 + k);
          osc_group(sa, Y0, Y1);
          osc_group(sb, Y0, Y1);
        }
      } else {
        sa.a0e = sa.a0o = sa.a1e = sa.a1o = make_float2(0.f, 0.f);
        sb.a0e = sb.a0o = sb.a1e = sb.a1o = make_float2(0.f, 0.f);
      }
      if (k < kc) {                               // kc % 4 != 0: uniform mask
        const float4 Y0 = mask4(*reinterpret_cast<const float4*>(x0 + k), k, kc);
        const float4 Y1 = mask4(*reinterpret_cast<const float4*>(x1 + k), k, kc);
        osc_group(sa, Y0, Y1);
        osc_group(sb, Y0, Y1);
      }
    } else {                 // live count changes inside this frame
      const int ka = live_harmonics(fa.x, fa.y, lc.fra, K, nyquist);
      const int kb = live_harmonics(fa.x, fa.y, lc.frb, K, nyquist);
      const int kmin = __reduce_min_sync(0xffffffffu, min(ka, kb));
      const int kmax = __reduce_max_sync(0xffffffffu, max(ka, kb));
      sa.a0e = sa.a0o = sa.a1e = sa.a1o = make_float2(0.f, 0.f);
      sb.a0e = sb.a0o = sb.a1e = sb.a1o = make_float2(0.f, 0.f);
      const int k_main = kmin & ~3;
      int k = 0;
      for (; k < k_main; k += 4) {
        const float4 Y0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 Y1 = *reinterpret_cast<const float4*>(x1 + k);
        osc_group(sa, Y0, Y1);
        osc_group(sb, Y0, Y1);
      }
      for (; k < kmax; k += 4) {
        const float4 Y0 = *reinterpret_cast<const float4*>(x0 + k);
        const float4 Y1 = *reinterpret_cast<const float4*>(x1 + k);
        osc_group(sa, mask4(Y0, k, ka), mask4(Y1, k, ka));
        osc_group(sb, mask4(Y0, k, kb), mask4(Y1, k, kb));
      }
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

#ifndef DDSP_HV3_MIN_CTAS
#define DDSP_HV3_MIN_CTAS 4
#endif
template <bool WINDOW, int HOPT>
__global__ void __launch_bounds__(NT, DDSP_HV3_MIN_CTAS)
harmonic_v3_kernel(HarmonicParams p, int use_tma, int FW) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hop = HOPT ? HOPT : p.hop;
  const int Kp = p.Kp, K = p.K, F = p.F;
  const int FT = FW * NW;
  const Smem L = smem_layout(FW, Kp, hop);
  void* mbar = (void*)(smem_raw + L.off_mbar);
  float2* sTab = (float2*)(smem_raw + L.off_tab);
  float* sX = (float*)(smem_raw + L.off_x);
  float* sW = (HOPT == 64) ? nullptr : (float*)(smem_raw + L.off_w);
  double* sRedD = (double*)(smem_raw + L.off_red);                        // [NW]
  unsigned long long* sWarpTot = (unsigned long long*)(smem_raw + L.off_red) + NW;

  const int b = blockIdx.y;
  const int i0 = blockIdx.x * FT;
  const int nfr = min(FT, F - i0);
  const int rows_in = min(nfr + 1, F - i0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* f0b = p.f0 + (size_t)b * F;
  const float* ampb = p.amps + (size_t)b * F;
  unsigned char* wbase = smem_raw + L.off_warp + warp * L.warp_stride;
  FrameRec* sRec = (FrameRec*)(wbase + L.w_rec);
  int* sLive = (int*)(wbase + L.w_live);

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

  // ---- 1. CTA-wide: sum of f0 over the frames before the tile (double), tables ----
  {
    double part = 0.0;
    for (int j = tid; j < i0; j += NT) part += (double)f0b[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) sRedD[warp] = part;
  }
  for (int j = tid; j < kSinTab; j += NT) sTab[j] = hcm::g_sincos256[j];
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
  unsigned long long excl = 0;            // wrapping sum of the warp's earlier frame totals
  {
    const int g0 = i0 + w0f;
    const int g = min(g0 + lane, F - 1);              // frame F := frame F-1
    float f = 0.f, a = 0.f;
    if (lane <= nfw && nfw > 0) {
      f = f0b[g];
      a = ampb[g];
      if (raw_scale) a = exp_sigmoid_f(a);            // synths.py:110-111
    }
    const float f_next = __shfl_down_sync(0xffffffffu, f, 1);
    const float a_next = __shfl_down_sync(0xffffffffu, a, 1);
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
    excl = incl - tot;
    if (lane == 31) sWarpTot[warp] = incl;             // total of the warp's frames
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
      FrameRec r;
      r.P = 0; r.A = Af; r.D = Df; r.kca = kca; r.kcb = kcb;
      r.f_lo = f; r.f_hi = f_next; r.amp0 = a; r.amp1 = a_next;
      sRec[lane] = r;
    }
    if (lane <= nfw && nfw > 0) sLive[lane] = live;
  }
  __syncthreads();            // tables, mbarrier init, partial sums, (LDG slab) visible

  // phase at the start of the tile (telescoped closed form, one double-precision
  // evaluation: <= 2^15 turns, 2^-38 turn resolution), then this warp's offset
  if (nfw > 0) {
    double base_sum = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) base_sum += sRedD[w];
    const double a_tile = (double)f0b[i0] * p.inv_sr;
    const double a_first = (double)f0b[0] * p.inv_sr;
    unsigned long long P0 = turns_to_fix64(
        (double)hop * (base_sum * p.inv_sr) + 0.5 * (hop - 1) * (a_tile - a_first));
    for (int w = 0; w < warp; ++w) P0 += sWarpTot[w];
    if (lane < nfw) sRec[lane].P = P0 + excl + 0x80000000ull;   // rounding offset folded in
  }
  if (use_tma) mbar_wait(mbar, 0);

  // ---- 3. get_controls on the warp's own rows (synths.py:110-117) ----
  const bool need_sync2 = have_ctl || (rows_in < nfr + 1);
  if (nfw > 0) {
    float* sXw = sX + (size_t)w0f * Kp;
    const bool last = (w0f + nfw == nfr);
    int nrows = nfw;
    if (last && rows_in > nfr) nrows = nfw + 1;         // the real row after the tile
    if (have_ctl) {
      if (Kp <= 128) {
        for (int r0 = 0; r0 < nrows; r0 += 4)
          controls_rows4(sXw, sLive, r0, nrows, Kp, raw_scale, lane);
      } else {
        for (int r0 = 0; r0 < nrows; r0 += 4)
          hcm::controls_rows(sXw, sLive, r0, nrows, Kp, raw_scale, lane);
      }
    }
    if (last && rows_in < nfr + 1) {                    // frame F := frame F-1
      __syncwarp();
      for (int c = lane; c < Kp; c += 32) sXw[nfw * Kp + c] = sXw[(nfw - 1) * Kp + c];
    }
  }
  if (need_sync2) __syncthreads();   // the row after a warp's block is its neighbour's
  else __syncwarp();                 // the warp's own frame records

  // ---- 4. samples ----
  if (nfw > 0) {
    float* outw = p.audio + (size_t)b * p.N + (size_t)(i0 + w0f) * hop;
    const float* xw = sX + (size_t)w0f * Kp;
    if (HOPT == 64) {
      const LaneConst lc = lane_const<WINDOW>(0, lane, inv_hop, nullptr);
      for (int li = 0; li < nfw; ++li) {
        frame_chunk(xw + li * Kp, xw + (li + 1) * Kp, sRec + li, lc, sTab, K,
                    p.nyquist, lane, outw + (size_t)li * 64, p.accumulate);
      }
    } else {
      for (int li = 0; li < nfw; ++li) {
        for (int r0 = 0; r0 < hop; r0 += 64) {
          const LaneConst lc = lane_const<WINDOW>(r0, lane, inv_hop, sW);
          frame_chunk(xw + li * Kp, xw + (li + 1) * Kp, sRec + li, lc, sTab, K,
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
    cudaError_t e = cudaFuncSetAttribute(harmonic_v3_kernel<WINDOW, HOPT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
  }
  harmonic_v3_kernel<WINDOW, HOPT><<<grid, NT, smem, st>>>(p, use_tma, FW);
  return cudaSuccess;
}

}  // namespace hv3

// Returns 0 on success, negative on error, 1 if the tile cannot fit shared memory
// (the caller then takes the generic kernel).
inline int launch_harmonic_v3(HarmonicParams p, cudaStream_t st) {
  using namespace hv3;
  p.Kp = (p.K + 3) & ~3;
  static const int env_fw = [] { const char* e = getenv("DDSP_B200_HARM_FW"); return e ? atoi(e) : 0; }();
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
    (void)cudaGetLastError();
    set_error("harmonic_forward(v3): %s", cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  DDSP_CHECK_LAUNCH("harmonic_forward(v3)");
  return 0;
}

}  // namespace ddsp
