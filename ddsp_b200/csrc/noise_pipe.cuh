// Warp-specialised, software-pipelined FilteredNoise kernel for the decoder
// shape (n_frequencies = 65, frame = 64 samples, 128-tap IR; ae.gin:60-68).
// Same maths and shared-memory formats as noise_fused.cuh (read that header
// first); what changes is WHO does what and WHEN:
//
//   * one persistent CTA per SM, 20 warps;
//   * warps 12..19 are PRODUCERS: for tile t+1 they stage the magnitudes
//     (cp.async), apply exp_sigmoid, synthesise the 32 impulse responses
//     (E/O cosine sums, lane = frame), write both tap copies and generate the
//     noise (Philox) - into stage (t+1) & 1 of a double buffer;
//   * warps 0..11 are CONSUMERS: for tile t they run the FFMA2 FIR out of stage
//     t & 1, overlap-add in shared memory and store the audio;
//   * the two groups meet only at two mbarriers per stage (full / empty), so
//     the FMA-bound FIR never waits for the latency-bound staging work.
//
// All geometry is compile-time here; every other shape takes noise_fused.cuh.
#pragma once
#include <string.h>

#include "noise_fused.cuh"
#include "noise_ring.cuh"

namespace ddsp {

namespace np_ {
constexpr int NB = 65, FRAME = 64, S = 128, S0 = 128, Q = 32, QP = 36;
constexpr int NE = 33, NO = 32, SHIFT = 64, START = 62, HB = 2, HA = 1, TFO = 29;
constexpr int NBLK = 12;                    // (FRAME + S - 1 + 15) / 16
constexpr int CONS_WARPS = 12, PROD_WARPS = 12, STAGES = 3;
constexpr int THREADS = 32 * (CONS_WARPS + PROD_WARPS);
constexpr int HS = 194, XS = 82, MS = 65;   // row strides (floats)
constexpr int PAD = kNfPad;                 // 32
constexpr int OUT_LEN = (FRAME + 1) * 35 + 17;   // even, keeps 8 B alignment
constexpr int NQ = FRAME / 4;               // 16 noise quads per frame

struct Smem {
  float te[NE * QP];
  float to[NO * QP];
  float win[S];
  float m[32 * MS + 2];
  alignas(16) float raw[32 * NB + 8];   // +8: 16-byte aligned bulk copies land with an offset
  float eo[4][2][32 * 9];    // [column block][E | O][lane * 9 + col] partial sums
  alignas(16) float h[STAGES][64 * HS];  // [stage][even copy 32 rows | odd copy 32 rows]
  alignas(16) float x[STAGES][32 * XS];
  float out[OUT_LEN];
  alignas(8) unsigned long long full[STAGES], empty[STAGES], rawbar;
};
}  // namespace np_

// One half (even-k or odd-k terms) of a block of 4*W4 columns of the cosine sums.
template <int W4>
__device__ __forceinline__ void ir_half(const float* __restrict__ mrow,
                                        const float* __restrict__ tab, int QPc,
                                        int nk, float* __restrict__ acc) {
#pragma unroll
  for (int c = 0; c < 4 * W4; ++c) acc[c] = 0.f;
#pragma unroll 3
  for (int k = 0; k < nk; ++k) {
    const float m = mrow[2 * k];
    const float4* t4 = reinterpret_cast<const float4*>(tab + k * QPc);
#pragma unroll
    for (int q = 0; q < W4; ++q) {
      const float4 c = t4[q];
      acc[4 * q + 0] = fmaf(m, c.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(m, c.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(m, c.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(m, c.w, acc[4 * q + 3]);
    }
  }
}

struct NoisePipeParams {
  const float* __restrict__ mags;
  const float* __restrict__ noise;
  float* audio;
  uint64_t seed, offset;
  int B, F, N, accumulate, raw;
  float bias;
  int tiles_per_item, n_tiles;
  int item_base;   // Philox item index of batch row 0 (chunked host pipeline)
};

__global__ void __launch_bounds__(np_::THREADS, 1)
noise_pipe_kernel(NoisePipeParams p) {
  using namespace np_;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float invS0 = 1.0f / (float)S0;

  // ---- once: tables, window, zero pads, barriers ----
  for (int e = tid; e < NE * QP; e += THREADS) {
    const int k = e / QP, n = e - k * QP;
    const int ph = (2 * k * n) % S0;
    const float ck = (k == 0 || 2 * k == NB - 1) ? invS0 : 2.0f * invS0;
    sm.te[e] = (n <= Q) ? ck * cospif(2.0f * (float)ph * invS0) : 0.f;
  }
  for (int e = tid; e < NO * QP; e += THREADS) {
    const int k = e / QP, n = e - k * QP;
    const int ph = ((2 * k + 1) * n) % S0;
    sm.to[e] = (n < Q) ? 2.0f * invS0 * cospif(2.0f * (float)ph * invS0) : 0.f;
  }
  for (int j = tid; j < S; j += THREADS)
    sm.win[j] = 0.5f - 0.5f * cospif(2.0f * (float)j / (float)S0);   // core.py:1498,1515
  for (int e = tid; e < STAGES * 64 * HS; e += THREADS) (&sm.h[0][0])[e] = 0.f;
  for (int e = tid; e < STAGES * 32 * XS; e += THREADS) (&sm.x[0][0])[e] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&sm.full[i], PROD_WARPS);
      mbar_init(&sm.empty[i], CONS_WARPS);
    }
    mbar_init(&sm.rawbar, 1);
  }
  __syncthreads();

  if (warp >= CONS_WARPS) {
    // =========================== PRODUCERS ===================================
    const int pw = warp - CONS_WARPS;           // 0..11
    const int ptid = tid - CONS_WARPS * 32;     // 0..383
    constexpr int PT = PROD_WARPS * 32;
    // Raw magnitudes of a tile -> sm.raw, asynchronously (used one tile later).
    // Interior tiles: ONE TMA bulk copy of the 16-byte aligned span that covers
    // the 32 contiguous rows (the rows are only 4-byte aligned: 65 floats each),
    // data lands at offset `roff`.  Edge tiles: per-element cp.async of the rows
    // that exist.  Completion: mbarrier `rawbar` (both paths arrive on it).
    int roff = 0;
    auto prefetch_mags = [&](int tile) -> int {
      const int b = tile / p.tiles_per_item;
      const int j0 = (tile - b * p.tiles_per_item) * TFO - HB;
      const float* magb = p.mags + (size_t)b * p.F * NB;
      const bool inside = (j0 >= 0) && (j0 + 32 < p.F || (j0 + 32 == p.F && b + 1 < p.B));
      if (inside) {
        const float* src = magb + (size_t)j0 * NB;
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const int off = (int)((a & 15) >> 2);
        if (ptid == 0) {
          const uint32_t bytes = (uint32_t)(((off + 32 * NB) * 4 + 15) & ~15);
          mbar_expect_tx(&sm.rawbar, bytes);
          tma_bulk_g2s(sm.raw, reinterpret_cast<const void*>(a & ~(uintptr_t)15),
                       bytes, &sm.rawbar);
        }
        return off;
      }
      for (int e = ptid; e < 32 * NB; e += PT) {
        const int j = j0 + e / NB;
        if (j >= 0 && j < p.F) cp_async4(sm.raw + e, magb + ((long long)j0 * NB + e));
      }
      cp_async_wait_all();
      named_bar(1, PT);
      if (ptid == 0) mbar_arrive(&sm.rawbar);
      return 0;
    };
    if ((int)blockIdx.x < p.n_tiles) roff = prefetch_mags(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int st = it % STAGES;
      const int b = tile / p.tiles_per_item;
      const int j0 = (tile - b * p.tiles_per_item) * TFO - HB;
      // A. magnitudes -> lane-private rows (exp_sigmoid fused)
      mbar_wait(&sm.rawbar, it & 1);
      named_bar(1, PT);            // every producer is done reading sm.m (previous tile)
      // columns 0..63: two full-warp passes per row; column 64: one lane per row
      for (int jl = pw; jl < 32; jl += PROD_WARPS) {
        const int j = j0 + jl;
        const float* src = sm.raw + roff + jl * NB;
        float* dst = sm.m + jl * MS;
        float v0 = 0.f, v1 = 0.f;
        if (j >= 0 && j < p.F) {
          v0 = src[lane];
          v1 = src[lane + 32];
          if (p.raw) {                                      // synths.py:176-177
            v0 = exp_sigmoid_f(v0 + p.bias);
            v1 = exp_sigmoid_f(v1 + p.bias);
          }
        }
        dst[lane] = v0;
        dst[lane + 32] = v1;
      }
      if (pw == PROD_WARPS - 1) {
        const int j = j0 + lane;
        float v = 0.f;
        if (j >= 0 && j < p.F) {
          v = sm.raw[roff + lane * NB + 64];
          if (p.raw) v = exp_sigmoid_f(v + p.bias);
        }
        sm.m[lane * MS + 64] = v;
      }
      named_bar(1, PT);            // sm.raw consumed, sm.m complete
      {
        const int nxt = tile + gridDim.x;
        if (nxt < p.n_tiles) roff = prefetch_mags(nxt);
      }
      // wait until the consumers have drained this stage (STAGES tiles ago)
      if (it >= STAGES) mbar_wait(&sm.empty[st], ((it / STAGES) - 1) & 1);
      float* hE = sm.h[st] + lane * HS + PAD;
      float* hO = sm.h[st] + (32 + lane) * HS + PAD;
      // B. IR synthesis: producer warps 0..3 sum the even-k terms (E) of column
      //    blocks 0..3, warps 4..7 the odd-k terms (O) of the same blocks; the
      //    pair swaps partial sums through shared memory and each writes half of
      //    the taps: E-warp h0[n] = E + O (offsets +-n), O-warp h0[64 - n] = E - O.
      if (pw < 8) {
        const int cb = pw & 3;                 // column block: n0 = 8 * cb
        const bool is_o = pw >= 4;
        const int n0 = cb * 8;
        float acc[12];
        const float* mrow = sm.m + lane * MS + (is_o ? 1 : 0);
        int ncol;
        if (is_o) {
          ir_half<2>(mrow, sm.to + n0, QP, NO, acc);
          ncol = 8;
        } else if (cb == 3) {                  // last E block also owns column 32
          ir_half<3>(mrow, sm.te + n0, QP, NE, acc);
          ncol = 9;
        } else {
          ir_half<2>(mrow, sm.te + n0, QP, NE, acc);
          ncol = 8;
        }
        float* mine = sm.eo[cb][is_o ? 1 : 0] + lane * 9;
        const float* other = sm.eo[cb][is_o ? 0 : 1] + lane * 9;
        if (is_o) acc[8] = 0.f;                // O[32] = 0
#pragma unroll
        for (int c = 0; c < 9; ++c) mine[c] = acc[c];
        named_bar(7 + cb, 64);                 // the E/O pair of this block
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          if (c < (cb == 3 ? 9 : 8)) {
            const int n = n0 + c;
            const float oth = other[c];
            if (!is_o) {
              const float hp = acc[c] + oth;   // |offset| = n
              int t = SHIFT + n;
              float v = sm.win[t & (S - 1)] * hp;
              if (t < S) { hE[t] = v; hO[t + 1] = v; }
              t = SHIFT - n;
              if (n != 0) { v = sm.win[t] * hp; hE[t] = v; hO[t + 1] = v; }
            } else if (n != Q) {
              const float hm = oth - acc[c];   // |offset| = 64 - n
              const int n2 = 2 * Q - n;
              int t = SHIFT + n2;
              if (t < S) { const float v = sm.win[t] * hm; hE[t] = v; hO[t + 1] = v; }
              t = SHIFT - n2;
              if (t >= 0) { const float v = sm.win[t] * hm; hE[t] = v; hO[t + 1] = v; }
            }
          }
        }
        (void)ncol;
      } else {
        // noise: 512 quads over producer warps 8..11
        const float* nzb = p.noise ? p.noise + (size_t)b * p.N : nullptr;
        const long long p_lo = (long long)j0 * FRAME;
        const bool interior = (p_lo >= 0) && (p_lo + 32ll * FRAME <= p.N) && !nzb;
        constexpr int TOTAL = 32 * NQ;             // 512 quads
        const int e_lo = (pw - 8) * (TOTAL / 4), e_hi = e_lo + TOTAL / 4;
        float* xs = sm.x[st];
        if (interior) {
          const uint32_t qbase = (uint32_t)(p_lo >> 2);
          for (int e = e_lo + lane; e < e_hi; e += 32) {
            const int jl = e >> 4, qd = e & 15;
            const float4 r = noise4(qbase + (uint32_t)e, (uint32_t)(b + p.item_base), p.seed, p.offset);
            float2* d = reinterpret_cast<float2*>(xs + jl * XS + 4 * qd);
            d[0] = make_float2(r.x, r.y);
            d[1] = make_float2(r.z, r.w);
          }
        } else {
          for (int e = e_lo + lane; e < e_hi; e += 32) {
            const int jl = e >> 4, qd = e & 15;
            const long long pp = (long long)(j0 + jl) * FRAME + 4 * qd;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (pp >= 0 && pp < p.N) {
              if (nzb) {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (pp + u < p.N) v[u] = nzb[pp + u];
              } else {
                const float4 r = noise4((uint32_t)(pp >> 2), (uint32_t)(b + p.item_base), p.seed,
                                        p.offset);
                v[0] = r.x;
                if (pp + 1 < p.N) v[1] = r.y;
                if (pp + 2 < p.N) v[2] = r.z;
                if (pp + 3 < p.N) v[3] = r.w;
              }
            }
            float* d = xs + jl * XS + 4 * qd;
            d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; d[3] = v[3];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[st]);
    }
  } else {
    // =========================== CONSUMERS ===================================
    constexpr int CT = CONS_WARPS * 32;        // 384
    const int n0 = warp * 16;                  // FIR block of this warp
    const int cls = warp & 3, grp = warp >> 2; // OLA class / frame group
    int it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int st = it % STAGES;
      const int b = tile / p.tiles_per_item;
      const int q0 = (tile - b * p.tiles_per_item) * TFO;
      float* outb = p.audio + (size_t)b * p.N;
      // += operand fetched early
      float4 pre[2];
      pre[0] = pre[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.accumulate) {
#pragma unroll
        for (int i2 = 0; i2 < 2; ++i2) {
          const int e = tid + i2 * CT;
          if (e < TFO * NQ) {
            const int ql = e >> 4, qd = e & 15;
            const int t = (q0 + ql) * FRAME + 4 * qd;
            if (q0 + ql < p.F && t + 3 < p.N &&
                (reinterpret_cast<uintptr_t>(outb + t) & 15) == 0)
              pre[i2] = *reinterpret_cast<const float4*>(outb + t);
          }
        }
      }
      mbar_wait(&sm.full[st], (it / STAGES) & 1);
      // D. FIR
      float2 acc2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc2[c] = make_float2(0.f, 0.f);
      {
        const float* xrow = sm.x[st] + lane * XS;
        const float* hE = sm.h[st] + lane * HS + PAD;
        const float* hO = sm.h[st] + (32 + lane) * HS + PAD;
        const int i_lo = max(0, n0 - (S - 1));
        const int i_hi = min(FRAME - 1, n0 + 15);
        const int ch_lo = i_lo >> 4, ch_hi = i_hi >> 4;
        const int bse = n0 - (ch_lo << 4);
        float2 WE[8], WO[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          WE[r] = *reinterpret_cast<const float2*>(hE + bse + 2 * r);
          WO[r] = *reinterpret_cast<const float2*>(hO + bse + 2 * r);
        }
        for (int ch = ch_lo; ch <= ch_hi; ++ch) {
          const int ib = ch << 4;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 xv = *reinterpret_cast<const float2*>(xrow + ib + 2 * e);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              acc2[c] = nf_ffma2(xv.x, WE[(c - e) & 7], acc2[c]);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              acc2[c] = nf_ffma2(xv.y, WO[(c - e) & 7], acc2[c]);
            const int nb2 = n0 - ib - 2 * e - 2;
            WE[(-e - 1) & 7] = *reinterpret_cast<const float2*>(hE + nb2);
            WO[(-e - 1) & 7] = *reinterpret_cast<const float2*>(hO + nb2);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[st]);   // stage may be refilled

      // E. overlap-add into the skewed buffer: slot = lane + grp, first writer
      //    stores, later ones add; the 3 warps of a class take turns.
      {
        float* orow = sm.out + (FRAME + 1) * (lane + grp) + cls * 16;
        const bool store = (grp == 0) || (lane == 31);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          if (grp == g) {
            if (store) {
#pragma unroll
              for (int c = 0; c < 8; ++c) { orow[2 * c] = acc2[c].x; orow[2 * c + 1] = acc2[c].y; }
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) { orow[2 * c] += acc2[c].x; orow[2 * c + 1] += acc2[c].y; }
            }
          }
          if (g < 2) named_bar(2 + cls, 96);
        }
      }
      named_bar(6, CT);
      // F. crop, (+= harmonic), store
#pragma unroll
      for (int i2 = 0; i2 < 2; ++i2) {
        const int e = tid + i2 * CT;
        if (e < TFO * NQ) {
          const int ql = e >> 4, qd = e & 15;
          const int t = (q0 + ql) * FRAME + 4 * qd;
          if (t < p.N && q0 + ql < p.F) {
            const int r0 = 4 * qd + START;           // < 2 * FRAME
            const int f0 = r0 >= FRAME ? 1 : 0;
            const int rem = r0 - f0 * FRAME;
            const int o = (ql + HB) * FRAME + r0;
            const int sk = ql + HB + f0;
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              v[u] = sm.out[o + u + sk + ((rem + u >= FRAME) ? 1 : 0)];
            if (t + 3 < p.N && (reinterpret_cast<uintptr_t>(outb + t) & 15) == 0) {
              float4 r = make_float4(v[0], v[1], v[2], v[3]);
              if (p.accumulate) {
                const float4 a = pre[i2];
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
              }
              *reinterpret_cast<float4*>(outb + t) = r;
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (t + u < p.N) {
                  float r = v[u];
                  if (p.accumulate) r += outb[t + u];
                  outb[t + u] = r;
                }
              }
            }
          }
        }
      }
      named_bar(6, CT);       // sOut is rewritten by the next tile's phase E
    }
  }
}

inline bool noise_pipe_supported(int F, int nb, int N, int window_size) {
  if (nb != np_::NB) return false;
  if (N % F != 0 || N / F != np_::FRAME) return false;
  IrGeom g = make_ir_geom(nb, window_size);
  return !g.padded && g.S == np_::S;
}

inline int launch_noise_pipe(const float* mags, const float* noise, uint64_t seed,
                             uint64_t offset, float* audio, int B, int F, int N,
                             int accumulate, cudaStream_t st, int raw, float bias,
                             int item_base = 0) {
  NoisePipeParams p;
  p.item_base = item_base;
  p.mags = mags; p.noise = noise; p.audio = audio; p.seed = seed; p.offset = offset;
  p.B = B; p.F = F; p.N = N; p.accumulate = accumulate; p.raw = raw; p.bias = bias;
  p.tiles_per_item = (F + np_::TFO - 1) / np_::TFO;
  const long long n_tiles = (long long)B * p.tiles_per_item;
  if (n_tiles >= (1ll << 31)) {
    set_error("filtered_noise_forward: too many tiles");
    return DDSP_B200_E_INVALID;
  }
  p.n_tiles = (int)n_tiles;
  const size_t smem = sizeof(np_::Smem);
  static_assert(sizeof(np_::Smem) <= 227 * 1024, "noise_pipe shared memory");
  cudaError_t e = cudaFuncSetAttribute(
      noise_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("filtered_noise_forward: cannot reserve %zu B smem: %s", smem,
              cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  const int grid = (int)std::min<long long>(n_tiles, (long long)kNumSMs);
  noise_pipe_kernel<<<grid, np_::THREADS, smem, st>>>(p);
  DDSP_CHECK_LAUNCH("filtered_noise_forward(pipelined)");
  return 0;
}

// Picks the pipelined kernel for the decoder shape, the generic fused one else.
inline int launch_noise_best(const float* mags, const float* noise, uint64_t seed,
                             uint64_t offset, float* audio, int B, int F, int nb,
                             int N, int window_size, int accumulate,
                             cudaStream_t st, int raw = 0, float bias = 0.f,
                            int item_base = 0, int overlap_previous = 0) {
  // noise_ring is the product kernel for the decoder shape; DDSP_B200_NOISE_IMPL
  // = pipe selects the second-generation kernel for A/B measurements.
  static const bool use_pipe = [] {
    const char* e = getenv("DDSP_B200_NOISE_IMPL");
    return e != nullptr && strcmp(e, "pipe") == 0;
  }();
  if (!use_pipe && noise_ring_supported(F, nb, N, window_size))
    return launch_noise_ring(mags, noise, seed, offset, audio, B, F, N, accumulate,
                             st, raw, bias, item_base, overlap_previous);
  if (noise_pipe_supported(F, nb, N, window_size))
    return launch_noise_pipe(mags, noise, seed, offset, audio, B, F, N, accumulate,
                             st, raw, bias, item_base);
  return launch_noise_fused(mags, noise, seed, offset, audio, B, F, nb, N,
                            window_size, accumulate, st, raw, bias, item_base);
}

}  // namespace ddsp
