// Warp-specialised, software-pipelined harmonic kernel for hop = 64 (every
// reference config).  Same maths as harmonic_fast.cuh - it reuses its per-frame
// pass, oscillator state and fused-controls code - but the per-tile prologue no
// longer stalls the oscillator warps:
//
//   * one persistent CTA per SM walks a CONTIGUOUS range of tiles (32 frames =
//     2048 samples each), so the phase at a tile start is carried from the
//     previous tile instead of being re-reduced from f0;
//   * warps 16..19 are PRODUCERS: for tile t+2 they issue the TMA bulk copy of
//     the harmonic_distribution slab into a 3-deep ring, stage f0 / amplitudes,
//     scan the frame phase totals (64-bit fixed point), compute the per-frame
//     live-harmonic counts and apply Harmonic.get_controls to the slab in place;
//   * warps 0..15 are CONSUMERS: two frames each per tile, the packed Reinsch
//     oscillator recurrence, coalesced stores;
//   * the groups meet only at mbarriers (full / empty per stage).
#pragma once
#include "harmonic_fast.cuh"

namespace ddsp {

namespace hp_ {
constexpr int HOP = 64, FT = 32;
constexpr int CONS_WARPS = 16, PROD_WARPS = 4, STAGES = 3;
constexpr int THREADS = 32 * (CONS_WARPS + PROD_WARPS);
constexpr int PT = 32 * PROD_WARPS;
constexpr int KMAX = 128;                     // slab row capacity (floats)

struct Stage {
  alignas(16) float x[(FT + 1) * KMAX];
  unsigned long long P[FT], A[FT], D[FT];
  float f0[FT + 2], amp[FT + 2];
  int kc[2 * FT];
  int nfr, i0, b, pad_;
};

struct Smem {
  Stage st[STAGES];
  float2 tab[kSinTab];
  float w[HOP];
  double red[2][PROD_WARPS];
  alignas(8) unsigned long long full[STAGES], empty[STAGES], tma[STAGES];
};
}  // namespace hp_

template <bool WINDOW>
__global__ void __launch_bounds__(hp_::THREADS, 1)
harmonic_pipe_kernel(HarmonicParams p, int tiles_per_item, int n_tiles,
                     int tiles_per_cta) {
  using namespace hp_;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = p.K, Kp = p.Kp, F = p.F;

  // ---- once: sin/cos table, interpolation weights, barriers ----
  for (int j = tid; j < kSinTab; j += THREADS) {
    float s, c;
    sincospif(2.0f * (float)j / (float)kSinTab, &s, &c);
    sm.tab[j] = make_float2(s, c);
  }
  for (int r = tid; r < HOP; r += THREADS) {
    const float frac = (float)r * (1.0f / (float)HOP);
    sm.w[r] = WINDOW ? (0.5f - 0.5f * cospif(frac)) : frac;
  }
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&sm.full[i], PROD_WARPS);
      mbar_init(&sm.empty[i], CONS_WARPS);
      mbar_init(&sm.tma[i], 1);
    }
  }
  __syncthreads();

  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(n_tiles, t_begin + tiles_per_cta);

  if (warp >= CONS_WARPS) {
    // =========================== PRODUCERS ===================================
    const int pw = warp - CONS_WARPS;
    const int ptid = tid - CONS_WARPS * 32;
    const bool raw_scale = p.ctl_flags & DDSP_B200_CTL_SCALE;
    const bool nyq = p.ctl_flags & DDSP_B200_CTL_NYQUIST;
    unsigned long long P_carry = 0;     // phase at the start of the next tile
    int carry_item = -1;                // item the carry belongs to
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int sidx = it % STAGES;
      Stage& S = sm.st[sidx];
      const int b = tile / tiles_per_item;
      const int i0 = (tile - b * tiles_per_item) * FT;
      const int nfr = min(FT, F - i0);
      const int rows_in = min(nfr + 1, F - i0);
      const float* f0b = p.f0 + (size_t)b * F;
      const float* ampb = p.amps + (size_t)b * F;
      if (it >= STAGES) mbar_wait(&sm.empty[sidx], ((it / STAGES) - 1) & 1);
      // 1. slab copy (TMA) - rows are 16-byte multiples here (K % 4 == 0)
      if (ptid == 0) {
        const uint32_t bytes = (uint32_t)rows_in * (uint32_t)K * 4u;
        mbar_expect_tx(&sm.tma[sidx], bytes);
        tma_bulk_g2s(S.x, p.hd + ((size_t)b * F + i0) * K, bytes, &sm.tma[sidx]);
        S.nfr = nfr; S.i0 = i0; S.b = b;
      }
      // 2. f0 / amplitudes of the tile's frames (+ the clamped frame after it)
      for (int j = ptid; j <= nfr; j += PT) {
        const int g = min(i0 + j, F - 1);
        S.f0[j] = f0b[g];
        const float a = ampb[g];
        S.amp[j] = raw_scale ? exp_sigmoid_f(a) : a;           // synths.py:110-111
      }
      // 3. phase at the tile start: carried, or reduced from f0 at an item change
      const bool need_prefix = (carry_item != b);
      if (need_prefix) {
        double part = 0.0;
        for (int j = ptid; j < i0; j += PT) part += (double)f0b[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) sm.red[it & 1][pw] = part;
      }
      named_bar(1, PT);
      // 4. warp 0: scan of frame phase totals; warps 1..3: live counts
      if (pw == 0) {
        unsigned long long P = P_carry;
        if (need_prefix) {
          double fsum = 0.0;
          for (int w = 0; w < PROD_WARPS; ++w) fsum += sm.red[it & 1][w];
          const double a_first = (double)f0b[0] * p.inv_sr;
          const double a_tile = (double)S.f0[0] * p.inv_sr;
          P = turns_to_fix64((double)HOP * (fsum * p.inv_sr) +
                             0.5 * (HOP - 1) * (a_tile - a_first));
        }
        const int j = lane;
        unsigned long long tot = 0;
        if (j < nfr) {
          const double a0 = (double)S.f0[j] * p.inv_sr;
          const double a1 = (double)S.f0[j + 1] * p.inv_sr;
          S.A[j] = turns_to_fix64(a0);
          S.D[j] = turns_to_fix64((a1 - a0) / (double)HOP);
          tot = turns_to_fix64((double)HOP * a0 + (a1 - a0) * (0.5 * (HOP - 1)));
        }
        unsigned long long incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += up;
        }
        if (j < nfr) S.P[j] = P + (incl - tot);
        P_carry = P + __shfl_sync(0xffffffffu, incl, 31);
      } else {
        for (int j = ptid - 32; j < nfr; j += PT - 32) {
          const float f_lo = S.f0[j], f_hi = S.f0[j + 1];
          if (f_lo >= 1.0f && f_hi >= 1.0f) {
            S.kc[2 * j] = live_harmonics(f_lo, f_hi, 0.0f, K, p.nyquist);
            S.kc[2 * j + 1] = live_harmonics(
                f_lo, f_hi, (float)(HOP - 1) * (1.0f / (float)HOP), K, p.nyquist);
          } else {
            S.kc[2 * j] = S.kc[2 * j + 1] = -1;
          }
        }
      }
      carry_item = b;      // (all producer threads track it; only warp 0 uses P_carry)
      // 5. the slab has landed: get_controls in place, then the clamped last row
      mbar_wait(&sm.tma[sidx], (it / STAGES) & 1);
      if (p.ctl_flags != 0) {
        for (int r0 = pw * 4; r0 < rows_in; r0 += PROD_WARPS * 4)
          harmonic_controls_rows(S.x, S.f0, r0, rows_in, K, Kp, p.nyquist, raw_scale,
                                 nyq, lane);
      }
      if (rows_in < nfr + 1) {
        named_bar(1, PT);
        for (int c = ptid; c < Kp; c += PT) S.x[nfr * Kp + c] = S.x[(nfr - 1) * Kp + c];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[sidx]);
    }
  } else {
    // =========================== CONSUMERS ===================================
    int it = 0;
    for (int tile = t_begin; tile < t_end; ++tile, ++it) {
      const int sidx = it % STAGES;
      Stage& S = sm.st[sidx];
      mbar_wait(&sm.full[sidx], (it / STAGES) & 1);
      const int nfr = S.nfr, i0 = S.i0, b = S.b;
      float* outb = p.audio + (size_t)b * p.N + (size_t)i0 * HOP;
      for (int li = warp; li < nfr; li += CONS_WARPS) {
        harmonic_frame_pass(S.x + li * Kp, S.x + (li + 1) * Kp, S.P[li], S.A[li],
                            S.D[li], S.kc[2 * li], S.kc[2 * li + 1], S.f0[li],
                            S.f0[li + 1], S.amp[li], S.amp[li + 1], sm.w, sm.tab, K,
                            p.nyquist, HOP, lane, outb + (size_t)li * HOP,
                            p.accumulate);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[sidx]);
    }
  }
}

inline bool harmonic_pipe_supported(const HarmonicParams& p) {
  return p.hop == hp_::HOP && p.hd != nullptr && (p.K % 4 == 0) &&
         p.K <= hp_::KMAX && (((uintptr_t)p.hd & 15) == 0);
}

// Returns 0 on success, negative on error.
inline int launch_harmonic_pipe(HarmonicParams p, cudaStream_t st) {
  using namespace hp_;
  p.Kp = p.K;
  p.FT = FT;
  const int tiles_per_item = (p.F + FT - 1) / FT;
  const long long n_tiles = (long long)p.B * tiles_per_item;
  if (n_tiles >= (1ll << 31)) {
    set_error("harmonic_forward: too many tiles");
    return DDSP_B200_E_INVALID;
  }
  const int grid = (int)std::min<long long>(n_tiles, (long long)kNumSMs);
  const int tiles_per_cta = (int)((n_tiles + grid - 1) / grid);
  const int grid2 = (int)((n_tiles + tiles_per_cta - 1) / tiles_per_cta);
  const size_t smem = sizeof(Smem);
  static_assert(sizeof(Smem) <= 227 * 1024, "harmonic_pipe shared memory");
  cudaError_t e;
  if (p.amp_method == DDSP_B200_AMP_WINDOW) {
    e = cudaFuncSetAttribute(harmonic_pipe_kernel<true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      harmonic_pipe_kernel<true><<<grid2, THREADS, smem, st>>>(p, tiles_per_item,
                                                            (int)n_tiles, tiles_per_cta);
  } else {
    e = cudaFuncSetAttribute(harmonic_pipe_kernel<false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      harmonic_pipe_kernel<false><<<grid2, THREADS, smem, st>>>(p, tiles_per_item,
                                                             (int)n_tiles, tiles_per_cta);
  }
  if (e != cudaSuccess) {
    set_error("harmonic_forward: cannot reserve %zu B smem: %s", smem,
              cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  DDSP_CHECK_LAUNCH("harmonic_forward(pipelined)");
  return 0;
}

}  // namespace ddsp
