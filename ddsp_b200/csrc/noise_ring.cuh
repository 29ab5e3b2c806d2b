// noise_ring: FilteredNoise.get_signal for the decoder shape (n_frequencies = 65,
// frame = 64 samples, 128-tap IR; ae.gin:60-68), third generation.  Same maths as
// noise_fused.cuh (windowed zero-phase IR per frame by E/O cosine
// sums, Philox noise, time-varying FIR == the reference's framed FFT convolution
// + overlap-add + crop, core.py:1382-1473); the second generation
// (profiles/experiments/noise_pipe.cuh.txt) was bound by
// shared-memory bandwidth (67 % of LSU wavefronts at 47 % FMA-pipe utilisation,
// profiles/r01_ncu_summary_v7.txt) and by consumer warps marching in lock step
// through an overlap-add buffer.  What changed:
//
//   * GATHER FORM, NO OVERLAP-ADD.  out[64 q + n] = sum_i x_q[i] h_q[n + 62 - i]
//     + sum_i x_{q+1}[i] h_{q+1}[n - 2 - i] + sum_i x_{q-1}[i] h_{q-1}[n + 126 - i]
//     (+ x_{q-2}[63] h_{q-2}[127] for n = 0), taps outside [0, 128) being zero:
//     exactly 128 MACs per output.  A lane owns output frame q and reads the rows
//     of frames q-2 .. q+1; a warp finishes its 32 x 32 output block in registers
//     and writes it straight to HBM (st / red.add for the fused Add) - consumer
//     warps never talk to each other.
//   * ROW RING, NO HALO.  A persistent CTA walks a contiguous range of frames; the
//     impulse responses and noise rows live in a ring of 32-row slots in shared
//     memory (lane-private rows, strides = 2 mod 4 floats: conflict-free LDS.64),
//     so neighbouring tiles share their edge rows instead of recomputing them.
//   * ONE WINDOW, TWO ACCUMULATOR SETS.  For an even input x[i] the tap pair
//     (h[m], h[m+1]) feeds the output pair (n, n+1); for the odd input x[i+1] the
//     SAME pair feeds (n+1, n+2).  Set A holds pairs (n, n+1), set B pairs
//     (n+1, n+2): every MAC is an FFMA2 on one 16-pair register window, one IR
//     copy in shared memory, 2 LDS.64 per 33 FFMA2 (was 3 per 16).
//   * TRIANGULAR TRIMMING at compile time: the rows of frames q+1 and q-1 cover
//     complementary triangles of the (n, i) square; fully unrolled bodies skip the
//     pairs whose taps are all out of range.
//   * producers: two groups of 4 warps alternate tiles; a group takes its tile
//     from raw magnitudes (TMA) through exp_sigmoid, both cosine half-sums (in
//     registers as FFMA2, no exchange) and the windowed taps to the Philox rows.
#pragma once
#include "noise_fused.cuh"

namespace ddsp {

namespace nr_ {
constexpr int NB = 65, FRAME = 64, S = 128, S0 = 128, Q = 32, QP = 36;
constexpr int NE = 33, NO = 32, SHIFT = 64;
// A consumer warp owns a UNIT of 2 NW outputs of 32 frames: NW tap pairs in its
// register window, NW + NW + 1 packed accumulators.  NW = 16 (two units per frame,
// ~150 registers, 9 KB FIR bodies) and NW = 8 (four units, 96 registers, bodies
// that fit the 6 KB L0 instruction cache, twice the shared-memory loads per
// FFMA2) were both measured: 343 vs 357 us per B=256 decoder step.  The FIR is
// bound by the FMA pipe either way - FFMA2 with a broadcast scalar operand issues
// every 2.4 cycles per sub-partition, not 2 (tools/microbench2.cu).
#ifndef DDSP_NR_NW
#define DDSP_NR_NW 16
#endif
constexpr int NW = DDSP_NR_NW;
constexpr int UPT = FRAME / (2 * NW);              // units per 64-sample frame
#ifndef DDSP_NR_PHILOX_EARLY
#define DDSP_NR_PHILOX_EARLY 0
#endif
#ifndef DDSP_NR_PROD_GROUPS
#define DDSP_NR_PROD_GROUPS 3
#endif
#ifndef DDSP_NR_MAX_SLOTS
#define DDSP_NR_MAX_SLOTS 7
#endif
constexpr int CONS_WARPS = 8, PROD_GROUPS = DDSP_NR_PROD_GROUPS, PROD_WARPS = 4 * PROD_GROUPS;
constexpr int NTG = CONS_WARPS / UPT;              // consumption tiles in flight
static_assert((NTG & (NTG - 1)) == 0, "tile groups: a power of two");
// 32-row slots in the ring: the consumers hold NTG + 1 of them, each producer
// group fills one more - as far as 227 KB of shared memory go (7 slots with three
// raw-magnitude staging buffers, 8 with two)
constexpr int SLOTS = DDSP_NR_MAX_SLOTS;
constexpr int RING = 32 * SLOTS;
constexpr int THREADS = 32 * (CONS_WARPS + PROD_WARPS);
// NW = 16 only: 640 threads launch with 96 registers each; the consumers (256
// threads) grow to CONS_REGS out of what the producers (384 threads) give back.
#ifndef DDSP_NR_PROD_REGS
#define DDSP_NR_PROD_REGS 64
#endif
#ifndef DDSP_NR_CONS_REGS
#define DDSP_NR_CONS_REGS 144
#endif
constexpr int CONS_REGS = (NW == 16) ? DDSP_NR_CONS_REGS : 0, PROD_REGS = DDSP_NR_PROD_REGS;
constexpr int HPAD = 2, HS = 134, XS = 66, MS = 65;   // row strides (floats)
constexpr int NQ = FRAME / 4;

// -DDDSP_NR_TIMING: per-warp cycle counters by phase (tools/noise_timing.py reads
// them back through ddsp_b200_debug_noise_timing); measurement builds only.
#ifdef DDSP_NR_TIMING
#define NR_TIMING_DECL unsigned tprev__ = (unsigned)clock()
#define NR_LAP(i)                                                        \
  do {                                                                   \
    const unsigned n__ = (unsigned)clock();                              \
    if (lane == 0) sm.tacc[warp * 8 + (i)] += n__ - tprev__;             \
    tprev__ = n__;                                                       \
  } while (0)
__device__ unsigned g_nr_timing[kNumSMs * 32 * 8];
#else
#define NR_TIMING_DECL
#define NR_LAP(i)
#endif

struct Smem {
#ifdef DDSP_NR_TIMING
  unsigned tacc[32 * 8];
#endif
  float te[NE * QP];
  float to[NO * QP];
  float win[S];
  alignas(16) float raw[PROD_GROUPS][32 * NB + 8];
  alignas(16) float h[RING * HS];
  alignas(16) float x[RING * XS];
  alignas(8) unsigned long long full[SLOTS], empty[SLOTS], rawbar[PROD_GROUPS];
};

struct Params {
  const float* __restrict__ mags;
  const float* __restrict__ noise;
  float* audio;
  uint64_t seed, offset;
  int B, F, N, accumulate, raw, item_base;
  float bias;
};

// A contiguous run of output frames [s0, s0 + len) of batch item b.
struct Seg {
  int b, s0, len, nP, nC;
};

__device__ __forceinline__ bool next_seg(long long& g, long long g1, int F, Seg& sg) {
  if (g >= g1) return false;
  sg.b = (int)(g / F);
  sg.s0 = (int)(g - (long long)sg.b * F);
  sg.len = (int)min((long long)(F - sg.s0), g1 - g);
  sg.nP = (sg.len + 3 + 31) >> 5;       // production tiles: rows s0-2 .. s0+len
  sg.nC = (sg.len + 31) >> 5;           // consumption tiles
  g += sg.len;
  return true;
}

__device__ __forceinline__ void mbar_arrive_n(void* bar, int n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(n)
               : "memory");
}

// Shared-memory loads the scheduler may neither merge nor reorder (software
// pipelines that need several loads in flight with distinct destinations).
__device__ __forceinline__ float4 lds128v(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ float lds32v(const float* p) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
  return v;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c,
                                           float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a),
               "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ---- consumer: FIR bodies ---------------------------------------------------
// Window convention: at step e (inputs 2e, 2e+1) W[(c - e) mod NW] = (h[m], h[m+1])
// with m = N0 + 2c + C - 2e: the tap pair that takes x[2e] to outputs (N0+2c,
// N0+2c+1) [set A] and x[2e+1] to (N0+2c+1, N0+2c+2) [set B].  Bm1 is set B's
// pair c = -1 (outputs N0-1, N0): at step e it takes x[2e-1] with W_e[0].
struct Acc {
  float2 A[NW], B[NW], Bm1;
};

// The FIR of one unit runs as a short PROGRAM of NW-step bodies, executed by a
// loop with a switch so that each body exists once in the instruction stream.
// A body walks NW steps k of one row: inputs xp[2k], xp[2k+1]; the tap pair of
// (c, k) sits at hp + 2c - 2k.
//   FULL : every pair (c, k) matters.
//   LOWER: only c >= k + 1 (taps left of the row start are zero), NW - 1 steps,
//          no new pair ever enters the window.
//   UPPER: only c <= k (taps right of the row end are zero); the window fills up
//          from pair 0.
enum { OP_FULL = 0, OP_LOWER = 1, OP_UPPER = 2 };
struct Op {
  const float* xp;
  const float* hp;
  int kind;
  int preload;      // 0: window continues, 1: load all NW pairs, 2: pair 0 only
  int set_xo;       // xo_prev := xo before the body
  float xo;
  int final_op;     // afterwards: Bm1 += xo_prev * W[0]  (the last odd input)
};
constexpr int N_OPS = 2 * UPT + 1;

// Op i of unit u (outputs N0 = 2 NW u ..).  x_m1 / h_m1: rows of frame q+1 (tap
// offset C = -2); x_0 / h_0: frame q (C = 62); x_p1 / h_p1: frame q-1 (C = 126).
// h_* point at tap 0.  Body j of a row covers inputs 2 NW j .. 2 NW (j+1) - 1.
//   frame q  : UPT FULL bodies (every (n, i) pair is in range); the closing op
//              (x[63] with taps (N0-2, N0-1)) is all padding for N0 = 0.
//   frame q+1: reaches outputs n >= i + 2: u FULL bodies, then the shrinking
//              triangle.
//   frame q-1: covers i >= n - 1: the growing triangle in body u (entered with
//              x[N0-1] -> output N0), then FULL bodies to the end of the row.
__device__ __forceinline__ Op block_op(int i, int u, const float* x_m1,
                                       const float* h_m1, const float* x_0,
                                       const float* h_0, const float* x_p1,
                                       const float* h_p1) {
  const int N0 = 2 * NW * u;
  if (i < UPT)
    return Op{x_0 + 2 * NW * i, h_0 + N0 + 62 - 2 * NW * i, OP_FULL, i == 0, i == 0,
              0.f, (i == UPT - 1) && (N0 != 0)};
  i -= UPT;
  if (i < u)
    return Op{x_m1 + 2 * NW * i, h_m1 + N0 - 2 - 2 * NW * i, OP_FULL, i == 0, i == 0,
              0.f, 0};
  if (i == u) return Op{x_m1 + N0, h_m1 - 2, OP_LOWER, u == 0, u == 0, 0.f, 0};
  i -= 1;                                              // body index in frame q-1
  if (i == u)
    return Op{x_p1 + N0, h_p1 + 126, OP_UPPER, 2, 1, (u > 0) ? x_p1[N0 - 1] : 0.f,
              u == UPT - 1};
  return Op{x_p1 + 2 * NW * i, h_p1 + N0 + 126 - 2 * NW * i, OP_FULL, 0, 0, 0.f,
            i == UPT - 1};
}

__device__ __forceinline__ void consume_block(Acc& a, int u, const float* x_m1,
                                              const float* h_m1, const float* x_0,
                                              const float* h_0, const float* x_p1,
                                              const float* h_p1) {
  float2 W[NW];
  float xo_prev = 0.f;
  constexpr int M = NW - 1;
#pragma unroll 1
  for (int i = 0; i < N_OPS; ++i) {
    const Op op = block_op(i, u, x_m1, h_m1, x_0, h_0, x_p1, h_p1);
    const float* __restrict__ xp = op.xp;
    const float* __restrict__ hp = op.hp;
    if (op.preload == 1) {
#pragma unroll
      for (int c = 0; c < NW; ++c) W[c] = *reinterpret_cast<const float2*>(hp + 2 * c);
    } else if (op.preload == 2) {
      W[0] = *reinterpret_cast<const float2*>(hp);
    }
    if (op.set_xo) xo_prev = op.xo;
    switch (op.kind) {
      case OP_FULL:
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          const float2 xv = *reinterpret_cast<const float2*>(xp + 2 * k);
          // last pair first: its slot is refilled (for step k + 1) right behind it
          a.A[M] = nf_ffma2(xv.x, W[(M - k) & M], a.A[M]);
          a.B[M] = nf_ffma2(xv.y, W[(M - k) & M], a.B[M]);
          const float2 wn = *reinterpret_cast<const float2*>(hp - 2 * (k + 1));
#pragma unroll
          for (int c = M - 1; c >= 0; --c) {
            a.A[c] = nf_ffma2(xv.x, W[(c - k) & M], a.A[c]);
            a.B[c] = nf_ffma2(xv.y, W[(c - k) & M], a.B[c]);
          }
          a.Bm1 = nf_ffma2(xo_prev, W[(0 - k) & M], a.Bm1);
          W[(M - k) & M] = wn;
          xo_prev = xv.y;
        }
        break;
      case OP_LOWER:
#pragma unroll
        for (int k = 0; k < NW - 1; ++k) {
          const float2 xv = *reinterpret_cast<const float2*>(xp + 2 * k);
#pragma unroll
          for (int c = M; c >= k + 1; --c) {
            a.A[c] = nf_ffma2(xv.x, W[(c - k) & M], a.A[c]);
            a.B[c] = nf_ffma2(xv.y, W[(c - k) & M], a.B[c]);
          }
          xo_prev = xv.y;
        }
        break;
      default:
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          const float2 xv = *reinterpret_cast<const float2*>(xp + 2 * k);
          const float2 wn = *reinterpret_cast<const float2*>(hp - 2 * (k + 1));
#pragma unroll
          for (int c = k; c >= 0; --c) {
            a.A[c] = nf_ffma2(xv.x, W[(c - k) & M], a.A[c]);
            a.B[c] = nf_ffma2(xv.y, W[(c - k) & M], a.B[c]);
          }
          a.Bm1 = nf_ffma2(xo_prev, W[(0 - k) & M], a.Bm1);
          W[(M - k) & M] = wn;
          xo_prev = xv.y;
        }
        break;
    }
    if (op.final_op) a.Bm1 = nf_ffma2(xo_prev, W[0], a.Bm1);
  }
}

__global__ void __launch_bounds__(nr_::THREADS, 1)
noise_ring_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float invS0 = 1.0f / (float)S0;

  // ---- once: tables, window, zeroed ring (pads stay zero), barriers ----
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
  for (int e = tid; e < RING * HS; e += THREADS) sm.h[e] = 0.f;
  for (int e = tid; e < RING * XS; e += THREADS) sm.x[e] = 0.f;
#ifdef DDSP_NR_TIMING
  for (int e = tid; e < 32 * 8; e += THREADS) sm.tacc[e] = 0;
#endif
  if (tid == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(&sm.full[i], 4);
      mbar_init(&sm.empty[i], 2 * UPT);
    }
    for (int i = 0; i < PROD_GROUPS; ++i) mbar_init(&sm.rawbar[i], 1);
  }
  __syncthreads();

  // this CTA's frames: an even share of the B * F frames, cut at item boundaries
  const long long T = (long long)p.B * p.F;
  const long long g_lo = T * blockIdx.x / gridDim.x;
  const long long g_hi = T * (blockIdx.x + 1) / gridDim.x;

  if (warp < CONS_WARPS) {
    // =========================== CONSUMERS ===================================
    if (CONS_REGS != 0)   // NW = 16: the FIR wants ~150 registers, from the producers
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONS_REGS ? CONS_REGS : 96));
    const int tg = warp / UPT, unit = warp % UPT;      // tile group, unit of the tile
    long long g = g_lo;
    Seg sg;
    int pbase = 0, ct = 0;                   // ct: consumption tiles before this segment
    bool dep_done = false;
    NR_TIMING_DECL;
    while (next_seg(g, g_hi, p.F, sg)) {
      // tiles ct + t of the CTA go round-robin over the tile groups
      for (int t = (tg - ct) & (NTG - 1); t < sg.nC; t += NTG) {
        NR_LAP(3);
        // rows of this tile live in production tiles t and t+1 of the segment
        // (two producer groups finish tiles out of order: wait for both)
        {
          const int P0 = pbase + t, P1 = pbase + min(t + 1, sg.nP - 1);
          mbar_wait(&sm.full[P0 % SLOTS], (P0 / SLOTS) & 1);
          mbar_wait(&sm.full[P1 % SLOTS], (P1 / SLOTS) & 1);
        }
        NR_LAP(0);
        const int q_rel = 32 * t + lane;                // output frame, relative
        // lanes past the end of the segment redo its last frame (and store
        // nothing): they must not wander into rows nobody produced
        const int rho = min(q_rel, sg.len - 1) + 2;     // its production row
        const int rbase = pbase * 32;
        auto xrow = [&](int r) { return sm.x + ((rbase + r) % RING) * XS; };
        auto hrow = [&](int r) { return sm.h + ((rbase + r) % RING) * HS + HPAD; };
        Acc a;
#pragma unroll
        for (int c = 0; c < NW; ++c) a.A[c] = a.B[c] = make_float2(0.f, 0.f);
        a.Bm1 = make_float2(0.f, 0.f);
        consume_block(a, unit, xrow(rho + 1), hrow(rho + 1), xrow(rho), hrow(rho),
                      xrow(rho - 1), hrow(rho - 1));
        // frame q-2 reaches output 0 only: input 63 through tap 127
        if (unit == 0)
          a.A[0].x = fmaf(xrow(rho - 2)[63], hrow(rho - 2)[127], a.A[0].x);
        __syncwarp();
        NR_LAP(1);
        if (lane == 0) {
          // release the slots: a production tile is read by consumption tiles
          // tau-1 and tau (UPT warps each); a lone reader arrives twice.
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const int tau = t + d;
            if (tau < sg.nP) {
              const int readers = (tau >= 1 ? 1 : 0) + (tau < sg.nC ? 1 : 0);
              mbar_arrive_n(&sm.empty[(pbase + tau) % SLOTS], 2 / readers);
            }
          }
        }
        // ---- store / accumulate this lane's 2 NW outputs ----
        if (!dep_done) {
          // programmatic dependent launch: this grid may have started while the
          // harmonic kernel was still draining; its audio must be complete (and
          // visible) before the first add lands.  No-op for a plain launch.
          asm volatile("griddepcontrol.wait;" ::: "memory");
          dep_done = true;
        }
        if (q_rel < sg.len) {
          constexpr int NOUT = 2 * NW;
          const long long t0 = (long long)(sg.s0 + q_rel) * FRAME + NOUT * unit;
          float* o = p.audio + (size_t)sg.b * p.N + t0;
          float v[NOUT];
#pragma unroll
          for (int c = 0; c < NW; ++c) {
            const float2 bp = (c == 0) ? a.Bm1 : a.B[c - 1];
            v[2 * c] = a.A[c].x + bp.y;
            v[2 * c + 1] = a.A[c].y + a.B[c].x;
          }
          if (t0 + NOUT <= p.N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
            for (int w4 = 0; w4 < NOUT / 4; ++w4) {
              if (p.accumulate)
                red_add_v4(o + 4 * w4, v[4 * w4], v[4 * w4 + 1], v[4 * w4 + 2],
                           v[4 * w4 + 3]);
              else
                *reinterpret_cast<float4*>(o + 4 * w4) =
                    make_float4(v[4 * w4], v[4 * w4 + 1], v[4 * w4 + 2], v[4 * w4 + 3]);
            }
          } else {
#pragma unroll
            for (int w1 = 0; w1 < NOUT; ++w1) {
              if (t0 + w1 < p.N) {
                if (p.accumulate) o[w1] += v[w1]; else o[w1] = v[w1];
              }
            }
          }
        }
        NR_LAP(2);
      }
      ct += sg.nC;
      pbase += sg.nP;
    }
  } else {
    // =============================== PRODUCERS ===============================
    if (CONS_REGS != 0)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PROD_REGS));
    // PROD_GROUPS groups of four warps; group g builds the production tiles P = g
    // (mod PROD_GROUPS) end to end: magnitudes (TMA) -> exp_sigmoid (in place) ->
    // both cosine half-sums -> windowed taps, then the tile's noise rows.  A group
    // has PROD_GROUPS tile periods for one tile, so its latency chains (TMA, MUFU,
    // LDS) stay off the consumers' critical path.
    const int grp = (warp - CONS_WARPS) >> 2;
    const int iw = (warp - CONS_WARPS) & 3;           // column block / slice
    const int ptid = tid - (CONS_WARPS + 4 * grp) * 32;   // 0..127 within the group
    constexpr int PT = 128;
    const int bar_id = 1 + grp;
    float* s_raw = sm.raw[grp];
    void* rawbar = &sm.rawbar[grp];
    const float* mags_end = p.mags + (size_t)p.B * p.F * NB;
    // Raw magnitudes of a production tile -> s_raw (asynchronously; consumed one
    // of the group's tiles later).  The valid rows are contiguous in HBM: ONE TMA
    // bulk copy of the enclosing 16-byte aligned span (rows are only 4-byte
    // aligned: 65 floats); row r lands at s_raw[roff + 65 r].  Spans that would
    // leave the tensor fall back to per-element cp.async.
    auto prefetch = [&](const Seg& s, int tau) -> int {
      const int jb = s.s0 - 2 + 32 * tau;
      const int r_lo = max(0, -jb), r_hi = min(32, p.F - jb);
      if (r_hi <= r_lo) {
        if (ptid == 0) mbar_arrive(rawbar);
        return 0;
      }
      const float* src = p.mags + ((size_t)s.b * p.F + jb + r_lo) * NB;
      const uintptr_t a = reinterpret_cast<uintptr_t>(src);
      const int off = (int)((a & 15) >> 2);
      const uint32_t bytes = (uint32_t)(((off + (r_hi - r_lo) * NB) * 4 + 15) & ~15);
      const uintptr_t a_al = a & ~(uintptr_t)15;
      if (a_al >= reinterpret_cast<uintptr_t>(p.mags) &&
          a_al + bytes <= reinterpret_cast<uintptr_t>(mags_end)) {
        if (ptid == 0) {
          mbar_expect_tx(rawbar, bytes);
          tma_bulk_g2s(s_raw, reinterpret_cast<const void*>(a_al), bytes, rawbar);
        }
        return off - r_lo * NB;
      }
      for (int e = ptid; e < (r_hi - r_lo) * NB; e += PT) cp_async4(s_raw + e, src + e);
      cp_async_wait_all();
      named_bar(bar_id, PT);
      if (ptid == 0) mbar_arrive(rawbar);
      return -r_lo * NB;
    };
    // iterator over the CTA's production tiles: (segment, tau), global index P
    struct It {
      long long g;     // frames consumed by next_seg so far
      Seg sg;
      int tau, P;
      bool ok;
    };
    auto it_begin = [&]() {
      It it;
      it.g = g_lo; it.tau = 0; it.P = 0;
      it.ok = next_seg(it.g, g_hi, p.F, it.sg);
      return it;
    };
    auto it_next = [&](It& it) {
      ++it.P;
      if (++it.tau >= it.sg.nP) {
        it.tau = 0;
        it.ok = next_seg(it.g, g_hi, p.F, it.sg);
      }
    };
    It cur = it_begin();
    for (int i = 0; i < grp && cur.ok; ++i) it_next(cur);
    int roff = 0;
    if (cur.ok) roff = prefetch(cur.sg, cur.tau);
    int n_mine = 0;                                   // tiles this group has staged
    NR_TIMING_DECL;
    while (cur.ok) {
      NR_LAP(7);
      It nxt = cur;
      for (int i = 0; i < PROD_GROUPS && nxt.ok; ++i) it_next(nxt);
      const Seg& sg = cur.sg;
      const int P = cur.P, slot = P % SLOTS;
      const int jb = sg.s0 - 2 + 32 * cur.tau;
      // A. exp_sigmoid in place on the raw rows (synths.py:176-177); rows of frames
      //    outside [0, F) hold nothing and are forced to zero taps below
      mbar_wait(rawbar, n_mine & 1);
      NR_LAP(0);
      const bool row_ok = (jb + lane >= 0) && (jb + lane < p.F);
      const float* mrow = row_ok ? s_raw + roff + lane * NB : s_raw;
      if (p.raw && row_ok) {
        float* src = s_raw + roff + lane * NB + iw * 17;
        float v[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) v[k] = (iw * 17 + k < NB) ? src[k] : 0.f;
#pragma unroll
        for (int k = 0; k < 17; ++k) v[k] = exp_sigmoid_f(v[k] + p.bias);
#pragma unroll
        for (int k = 0; k < 17; ++k)
          if (iw * 17 + k < NB) src[k] = v[k];
      }
      named_bar(bar_id, PT);       // rows complete
      ++n_mine;
      NR_LAP(1);
#ifdef DDSP_NR_EARLY_EMPTY_WAIT
      if (P >= SLOTS) mbar_wait(&sm.empty[slot], ((P / SLOTS) - 1) & 1);
#endif
      // the tile's noise rows (C. below): 512 quads, 128 per warp
      const float* nzb = p.noise ? p.noise + (size_t)sg.b * p.N : nullptr;
      const uint32_t item = (uint32_t)(sg.b + p.item_base);
      const long long p_lo = (long long)jb * FRAME;
      const bool interior = (jb >= 0) && (jb + 32 <= p.F) &&
                            (p_lo + 32ll * FRAME <= p.N) && !nzb;
      constexpr int PER = 32 * NQ / 4;                 // 128 quads per warp
#if DDSP_NR_PHILOX_EARLY
      float4 nzq[PER / 32];
#endif
      // B. both half-size cosine sums of 8 columns (9 for the last block):
      //    h0[n] = E[n] + O[n], h0[64 - n] = E[n] - O[n]   
      {
        const int n0 = 8 * iw;
        float2 aE[4], aO[4];
        float e8 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) aE[c] = aO[c] = make_float2(0.f, 0.f);
        // Software-pipelined over k, two register sets: the loads of term k + 1
        // are in flight while term k is multiplied.  (Left to ptxas, every
        // LDS.128 of this loop landed in the same four registers - one load in
        // flight per warp, the loop ran at shared-memory latency: 35 % of the
        // producers' stall samples were short-scoreboard waits.)  The loads are
        // volatile asm so that their order and their distinct destinations stay.
        struct Term { float me, mo; float4 e0, e1, o0, o1; };
        auto load_term = [&](Term& t, int k) {
          const float* pm = mrow + 2 * k;
          const float* pe = sm.te + k * QP + n0;
          const float* po = sm.to + k * QP + n0;
          t.me = lds32v(pm);
          t.mo = lds32v(pm + 1);
          t.e0 = lds128v(pe); t.e1 = lds128v(pe + 4);
          t.o0 = lds128v(po); t.o1 = lds128v(po + 4);
        };
        auto mac_term = [&](const Term& t, int k) {
          aE[0] = nf_ffma2(t.me, make_float2(t.e0.x, t.e0.y), aE[0]);
          aE[1] = nf_ffma2(t.me, make_float2(t.e0.z, t.e0.w), aE[1]);
          aE[2] = nf_ffma2(t.me, make_float2(t.e1.x, t.e1.y), aE[2]);
          aE[3] = nf_ffma2(t.me, make_float2(t.e1.z, t.e1.w), aE[3]);
          aO[0] = nf_ffma2(t.mo, make_float2(t.o0.x, t.o0.y), aO[0]);
          aO[1] = nf_ffma2(t.mo, make_float2(t.o0.z, t.o0.w), aO[1]);
          aO[2] = nf_ffma2(t.mo, make_float2(t.o1.x, t.o1.y), aO[2]);
          aO[3] = nf_ffma2(t.mo, make_float2(t.o1.z, t.o1.w), aO[3]);
          if (iw == 3) e8 = fmaf(t.me, sm.te[k * QP + Q], e8);
        };
        Term tA, tB;
        load_term(tA, 0);
#pragma unroll 1
        for (int k = 0; k < NO; k += 2) {
          load_term(tB, k + 1);
          mac_term(tA, k);
          // k + 2 = 32 is the last, even-only term: its odd operands are read
          // (row 32 of `to` runs into `win`, mrow[65] into the next row / the pad)
          // and never used
          load_term(tA, k + 2);
          mac_term(tB, k + 1);
        }
        {
          aE[0] = nf_ffma2(tA.me, make_float2(tA.e0.x, tA.e0.y), aE[0]);
          aE[1] = nf_ffma2(tA.me, make_float2(tA.e0.z, tA.e0.w), aE[1]);
          aE[2] = nf_ffma2(tA.me, make_float2(tA.e1.x, tA.e1.y), aE[2]);
          aE[3] = nf_ffma2(tA.me, make_float2(tA.e1.z, tA.e1.w), aE[3]);
          if (iw == 3) e8 = fmaf(tA.me, sm.te[NO * QP + Q], e8);
        }
        NR_LAP(3);
        named_bar(bar_id, PT);     // the group is done reading the rows
        if (nxt.ok) roff = prefetch(nxt.sg, nxt.tau);
        NR_LAP(4);
        // Only now does the group need its ring slot: the cosine sums above live in
        // registers.  (The wait used to sit in front of them: with two free slots for
        // three groups, a tile took ~8300 cycles from "slot free" to "full" - 1.7 tile
        // periods - and consumers waited on `full` 8.6 % of their time while producers
        // waited on `empty` 32 % of theirs; tools/noise_timing.py.)
#if DDSP_NR_PHILOX_EARLY
        // interior tiles: the noise quads are drawn into registers as well, so that
        // all that is left to do once the slot is free are stores
        if (interior) {
          const uint32_t qbase = (uint32_t)(p_lo >> 2);
#pragma unroll
          for (int it = 0; it < PER / 32; ++it)
            nzq[it] = noise4(qbase + (uint32_t)(iw * PER + it * 32 + lane), item, p.seed, p.offset);
        }
#endif
#ifndef DDSP_NR_EARLY_EMPTY_WAIT
        if (P >= SLOTS) mbar_wait(&sm.empty[slot], ((P / SLOTS) - 1) & 1);
#endif
        NR_LAP(2);
        float* hr = sm.h + (slot * 32 + lane) * HS + HPAD;
        const float E[8] = {aE[0].x, aE[0].y, aE[1].x, aE[1].y,
                            aE[2].x, aE[2].y, aE[3].x, aE[3].y};
        const float O[8] = {aO[0].x, aO[0].y, aO[1].x, aO[1].y,
                            aO[2].x, aO[2].y, aO[3].x, aO[3].y};
        // window: two 16-byte loads per half (the periodic Hann window is symmetric
        // about tap 64: win[64 - n] == win[64 + n], win[128 - n] == win[n]); taps
        // leave as 8-byte stores where the pair is this lane's own
        const float4 wp0 = *reinterpret_cast<const float4*>(sm.win + SHIFT + n0);
        const float4 wp1 = *reinterpret_cast<const float4*>(sm.win + SHIFT + n0 + 4);
        const float4 wm0 = *reinterpret_cast<const float4*>(sm.win + n0);
        const float4 wm1 = *reinterpret_cast<const float4*>(sm.win + n0 + 4);
        const float WP[8] = {wp0.x, wp0.y, wp0.z, wp0.w, wp1.x, wp1.y, wp1.z, wp1.w};
        const float WM[8] = {wm0.x, wm0.y, wm0.z, wm0.w, wm1.x, wm1.y, wm1.z, wm1.w};
        float vp[8], vm[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          vp[c] = row_ok ? WP[c] * (E[c] + O[c]) : 0.f;   // |offset| = n
          vm[c] = row_ok ? WM[c] * (E[c] - O[c]) : 0.f;   // |offset| = 64 - n
        }
#pragma unroll
        for (int c = 0; c < 8; c += 2) {                 // taps 64 + n and 64 - (64 - n)
          *reinterpret_cast<float2*>(hr + SHIFT + n0 + c) = make_float2(vp[c], vp[c + 1]);
          *reinterpret_cast<float2*>(hr + n0 + c) = make_float2(vm[c], vm[c + 1]);
        }
#pragma unroll
        for (int c = 1; c < 7; c += 2) {                 // taps 64 - n and 128 - n, descending
          *reinterpret_cast<float2*>(hr + SHIFT - n0 - c - 1) = make_float2(vp[c + 1], vp[c]);
          *reinterpret_cast<float2*>(hr + S - n0 - c - 1) = make_float2(vm[c + 1], vm[c]);
        }
        hr[SHIFT - n0 - 7] = vp[7];
        hr[S - n0 - 7] = vm[7];
        if (n0 != 0) {
          hr[SHIFT - n0] = vp[0];
          hr[S - n0] = vm[0];                            // tap 64 + (64 - n)
        }
        if (iw == 3) {                                 // n = 32: O[32] = 0
          if (!row_ok) e8 = 0.f;
          hr[SHIFT + Q] = sm.win[SHIFT + Q] * e8;
          hr[SHIFT - Q] = sm.win[SHIFT - Q] * e8;
        }
      }
      NR_LAP(5);
      // C. the tile's noise rows
      {
        float* xs = sm.x + slot * 32 * XS;
        if (interior) {
#if !DDSP_NR_PHILOX_EARLY
          const uint32_t qbase = (uint32_t)(p_lo >> 2);
#endif
#pragma unroll
          for (int it = 0; it < PER / 32; ++it) {
            const int e = iw * PER + it * 32 + lane;
            const int r = e >> 4, qd = e & 15;
#if DDSP_NR_PHILOX_EARLY
            const float4 v = nzq[it];
#else
            const float4 v = noise4(qbase + (uint32_t)e, item, p.seed, p.offset);
#endif
            float2* d = reinterpret_cast<float2*>(xs + r * XS + 4 * qd);
            d[0] = make_float2(v.x, v.y);
            d[1] = make_float2(v.z, v.w);
          }
        } else {
          for (int it = 0; it < PER / 32; ++it) {
            const int e = iw * PER + it * 32 + lane;
            const int r = e >> 4, qd = e & 15;
            const int j = jb + r;
            const long long pp = (long long)j * FRAME + 4 * qd;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (j >= 0 && j < p.F && pp < p.N) {
              if (nzb) {
#pragma unroll
                for (int u = 0; u < 4; ++u) if (pp + u < p.N) v[u] = nzb[pp + u];
              } else {
                const float4 r4 = noise4((uint32_t)(pp >> 2), item, p.seed, p.offset);
                v[0] = r4.x;
                if (pp + 1 < p.N) v[1] = r4.y;
                if (pp + 2 < p.N) v[2] = r4.z;
                if (pp + 3 < p.N) v[3] = r4.w;
              }
            }
            float2* d = reinterpret_cast<float2*>(xs + r * XS + 4 * qd);
            d[0] = make_float2(v[0], v[1]);
            d[1] = make_float2(v[2], v[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[slot]);
      NR_LAP(6);
      cur = nxt;
    }
  }
#ifdef DDSP_NR_TIMING
  __syncthreads();
  for (int e = tid; e < 32 * 8; e += THREADS)
    g_nr_timing[blockIdx.x * 32 * 8 + e] = sm.tacc[e];
#endif
}

}  // namespace nr_

inline bool noise_ring_supported(int F, int nb, int N, int window_size) {
  if (nb != nr_::NB) return false;
  if (N % F != 0 || N / F != nr_::FRAME) return false;
  IrGeom g = make_ir_geom(nb, window_size);
  return !g.padded && g.S == nr_::S;
}

inline int launch_noise_ring(const float* mags, const float* noise, uint64_t seed,
                             uint64_t offset, float* audio, int B, int F, int N,
                             int accumulate, cudaStream_t st, int raw, float bias,
                             int item_base, int overlap_previous = 0) {
  nr_::Params p;
  p.mags = mags; p.noise = noise; p.audio = audio; p.seed = seed; p.offset = offset;
  p.B = B; p.F = F; p.N = N; p.accumulate = accumulate; p.raw = raw; p.bias = bias;
  p.item_base = item_base;
  const long long T = (long long)B * F;
  const size_t smem = sizeof(nr_::Smem);
  static_assert(sizeof(nr_::Smem) <= 227 * 1024, "noise_ring shared memory");
  cudaError_t e = cudaFuncSetAttribute(
      nr_::noise_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("filtered_noise_forward: cannot reserve %zu B smem: %s", smem,
              cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  // one persistent CTA per SM; tiny workloads get one CTA per 32-frame tile
  const int grid = (int)std::max<long long>(
      1, std::min<long long>((long long)kNumSMs, (T + 31) / 32));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(nr_::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  // Programmatic dependent launch (decoder path): let this grid's CTAs start on
  // SMs the harmonic kernel has already vacated - tables, TMA, the first impulse
  // responses and noise rows do not depend on it; the consumers wait
  // (griddepcontrol.wait) before their first add into the audio buffer.
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = overlap_previous ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, nr_::noise_ring_kernel, p);
  if (e != cudaSuccess) {
    set_error("filtered_noise_forward(ring): launch failed: %s", cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  DDSP_CHECK_LAUNCH("filtered_noise_forward(ring)");
  return 0;
}

// noise_ring for the decoder shape (65 bands, 64-sample frames, 128 taps), the
// generic fused kernel (noise_fused.cuh) for every other shape it supports.
inline int launch_noise_best(const float* mags, const float* noise, uint64_t seed,
                             uint64_t offset, float* audio, int B, int F, int nb,
                             int N, int window_size, int accumulate,
                             cudaStream_t st, int raw = 0, float bias = 0.f,
                             int item_base = 0, int overlap_previous = 0) {
  if (noise_ring_supported(F, nb, N, window_size))
    return launch_noise_ring(mags, noise, seed, offset, audio, B, F, N, accumulate,
                             st, raw, bias, item_base, overlap_previous);
  return launch_noise_fused(mags, noise, seed, offset, audio, B, F, nb, N,
                            window_size, accumulate, st, raw, bias, item_base);
}

}  // namespace ddsp
