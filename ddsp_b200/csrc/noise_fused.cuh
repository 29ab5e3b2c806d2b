// Fused FilteredNoise.get_signal (synths.py:181-196): magnitudes -> windowed
// impulse responses (core.py:1534-1565, 1477-1531) -> time-varying FIR of
// uniform noise (core.py:1382-1473) [-> += harmonic audio, processors.py:174].
// IRs, noise and per-frame partial convolutions live only in shared memory.
//
// Formulation.  The reference frames the input, convolves
// frame j with ITS impulse response h_j (FFT size >= frame + S - 1, i.e. a
// linear convolution) and overlap-adds:
//     y_j[n]  = sum_i x_j[i] h_j[n - i],  n in [0, frame + S - 1)
//     out[t]  = sum_j y_j[t + start - j * frame]
// Mapping: one LANE owns one input frame j (x_j and h_j are private rows in
// shared memory, odd strides -> no bank conflicts), one WARP owns a block of 16
// outputs n.  The inner loop slides a 16-tap register window over h_j: per
// input sample 1 LDS (x) + 1 LDS (new tap) feed 16 FFMAs.  A persistent grid
// (one resident wave) amortises the cosine table over many tiles.
#pragma once
#include "noise.cuh"

namespace ddsp {

constexpr int kNfThreads = 384;          // 12 warps
constexpr int kNfWarps = kNfThreads / 32;
constexpr int kNfR = 16;                 // outputs per thread in the FIR
constexpr int kNfPad = 32;               // zero taps either side of h rows
constexpr int kNfMaxNb = 80;             // cos table [nb][nb] must fit

struct NoiseFusedParams {
  const float* __restrict__ mags;   // [B,F,nb]
  const float* __restrict__ noise;  // [B,N] or nullptr
  float* audio;                     // [B,N]
  uint64_t seed, offset;
  int B, F, nb, N, frame, start, accumulate;
  int raw;                          // mags are raw network outputs:
  float bias;                       //   exp_sigmoid(x + bias) while staging
  int TFo, Hb, Ha;                  // output frames per tile, halo before/after
  int tiles_per_item, n_tiles;
  int mS, hS, xS, yS;               // smem row strides (floats)
  int ylen;                         // frame + S - 1
  IrGeom g;
};

struct NfSmem {
  size_t off_cos, off_win, off_m, off_h, off_x, off_y, total;
};

__host__ __device__ inline NfSmem nf_smem_layout(const NoiseFusedParams& p) {
  NfSmem s;
  size_t o = 0;
  const int nhp = (p.g.S0 / 2 + 1 + 3) & ~3;        // |n| values, padded to 4
  s.off_cos = o; o += sizeof(float) * (size_t)p.nb * nhp;  // [k][n]
  s.off_win = o; o += sizeof(float) * (size_t)p.g.S;
  s.off_m = o;   o += sizeof(float) * 32 * (size_t)p.mS;
  s.off_h = o;   o += sizeof(float) * 32 * (size_t)p.hS;
  s.off_x = o;   o += sizeof(float) * 32 * (size_t)p.xS;
  // y aliases the cos/m region?  kept separate in v1 for clarity.
  s.off_y = o;   o += sizeof(float) * 32 * (size_t)p.yS;
  s.total = (o + 15) & ~(size_t)15;
  return s;
}

__global__ void __launch_bounds__(kNfThreads, 2)
noise_fused_kernel(NoiseFusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const NfSmem L = nf_smem_layout(p);
  float* sCos = (float*)(smem_raw + L.off_cos);
  float* sWin = (float*)(smem_raw + L.off_win);
  float* sM = (float*)(smem_raw + L.off_m);
  float* sH = (float*)(smem_raw + L.off_h);
  float* sX = (float*)(smem_raw + L.off_x);
  float* sY = (float*)(smem_raw + L.off_y);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const IrGeom g = p.g;
  const int nb = p.nb, S = g.S, S0 = g.S0, frame = p.frame;
  const int nh = S0 / 2 + 1;
  const int nhp = (nh + 3) & ~3;

  // ---- once per CTA: cos table [k][n] (coefficient c_k / S0 folded in) and
  //      the causal window ----
  {
    const float inv = 1.0f / (float)S0;
    for (int e = tid; e < nb * nhp; e += kNfThreads) {
      const int k = e / nhp, n = e - k * nhp;
      const int ph = (int)(((long long)k * n) % S0);
      const float c = (k == 0 || k == nb - 1) ? inv : 2.0f * inv;
      sCos[e] = (n < nh) ? c * cospif(2.0f * (float)ph / (float)S0) : 0.f;
    }
    for (int j = tid; j < S; j += kNfThreads) {
      int idx; float w;
      ir_tap(g, j, &idx, &w);
      sWin[j] = w;
    }
    // zero the h pads once (taps are rewritten every tile, pads never)
    for (int e = tid; e < 32 * p.hS; e += kNfThreads) sH[e] = 0.f;
    for (int e = tid; e < 32 * p.xS; e += kNfThreads) sX[e] = 0.f;
  }
  __syncthreads();

  const int NJ = p.TFo + p.Hb + p.Ha;            // input frames per tile (<= 32)
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_item;
    const int q0 = (tile - b * p.tiles_per_item) * p.TFo;   // first output frame
    const int j0 = q0 - p.Hb;                                // first input frame
    const float* magb = p.mags + (size_t)b * p.F * nb;

    // ---- 1. stage magnitudes and noise for input frames j0 .. j0+NJ-1 ----
    for (int e = tid; e < NJ * nb; e += kNfThreads) {
      const int jl = e / nb, k = e - jl * nb;
      const int j = j0 + jl;
      float m = 0.f;
      if (j >= 0 && j < p.F) {
        m = magb[(size_t)j * nb + k];
        if (p.raw) m = exp_sigmoid_f(m + p.bias);   // synths.py:176-177
      }
      sM[jl * p.mS + k] = m;
    }
    {
      const long long p_lo = (long long)j0 * frame;
      const long long p_hi = (long long)(j0 + NJ) * frame;   // exclusive
      const long long q_lo = (p_lo >= 0 ? p_lo : 0) >> 2;
      const long long q_hi = ((p_hi < p.N ? p_hi : p.N) + 3) >> 2;
      // zero everything first when the tile touches the signal edges
      if (p_lo < 0 || p_hi > p.N) {
        for (int e = tid; e < NJ * p.xS; e += kNfThreads) sX[e] = 0.f;
        __syncthreads();
      }
      const float* nzb = p.noise ? p.noise + (size_t)b * p.N : nullptr;
      for (long long q = q_lo + tid; q < q_hi; q += kNfThreads) {
        float v[4];
        if (nzb) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const long long pp = 4 * q + u;
            v[u] = pp < p.N ? nzb[pp] : 0.f;
          }
        } else {
          const float4 r = noise4((uint32_t)q, (uint32_t)b, p.seed, p.offset);
          v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const long long pp = 4 * q + u;
          if (pp >= p_lo && pp < p_hi && pp < p.N) {
            const int rel = (int)(pp - p_lo);
            const int jl = rel / frame, i = rel - jl * frame;
            sX[jl * p.xS + i] = v[u];
          }
        }
      }
    }
    __syncthreads();

    // ---- 2. impulse responses: lane = frame, warp = block of 8 |n| values ----
    //   h0[n] = sum_k (c_k/S0) M_k cos(2 pi k n / S0), n = 0 .. S0/2
    //   tap j <-> zero-phase offset nz = j - shift; h[j] = win[j] * h0[|nz|]
    for (int n0 = warp * 8; n0 < nh; n0 += kNfWarps * 8) {
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      const float* mrow = sM + lane * p.mS;
      if (n0 + 4 < nhp) {
#pragma unroll 5
        for (int k = 0; k < nb; ++k) {
          const float m = mrow[k];
          const float4 ca = *reinterpret_cast<const float4*>(sCos + k * nhp + n0);
          const float4 cb = *reinterpret_cast<const float4*>(sCos + k * nhp + n0 + 4);
          acc[0] = fmaf(m, ca.x, acc[0]); acc[1] = fmaf(m, ca.y, acc[1]);
          acc[2] = fmaf(m, ca.z, acc[2]); acc[3] = fmaf(m, ca.w, acc[3]);
          acc[4] = fmaf(m, cb.x, acc[4]); acc[5] = fmaf(m, cb.y, acc[5]);
          acc[6] = fmaf(m, cb.z, acc[6]); acc[7] = fmaf(m, cb.w, acc[7]);
        }
      } else {
        for (int k = 0; k < nb; ++k) {
          const float m = mrow[k];
          const float4 ca = *reinterpret_cast<const float4*>(sCos + k * nhp + n0);
          acc[0] = fmaf(m, ca.x, acc[0]); acc[1] = fmaf(m, ca.y, acc[1]);
          acc[2] = fmaf(m, ca.z, acc[2]); acc[3] = fmaf(m, ca.w, acc[3]);
        }
      }
      float* hrow = sH + lane * p.hS + kNfPad;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int n = n0 + c;
        if (n < nh) {
          // taps whose |zero-phase offset| (mod S0) equals n
          const int ja = g.shift + n, jb = g.shift - n;
          if (ja < S) hrow[ja] = sWin[ja] * acc[c];
          if (jb >= 0 && jb < S && jb != ja) hrow[jb] = sWin[jb] * acc[c];
        }
      }
    }
    __syncthreads();

    // ---- 3. FIR: lane = frame j, warp = 16-output block of y_j ----
    {
      const float* xrow = sX + lane * p.xS;
      const float* hrow = sH + lane * p.hS + kNfPad;
      float* yrow = sY + lane * p.yS;
      const int nblk = (p.ylen + kNfR - 1) / kNfR;
      const int nchunk = (frame + 15) >> 4;
      for (int blk = warp; blk < nblk; blk += kNfWarps) {
        const int n0 = blk * kNfR;
        float acc[kNfR];
#pragma unroll
        for (int c = 0; c < kNfR; ++c) acc[c] = 0.f;
        const int i_lo = max(0, n0 - (S - 1));
        const int i_hi = min(frame - 1, n0 + kNfR - 1);
        for (int ch = i_lo >> 4; ch <= (i_hi >> 4) && ch < nchunk; ++ch) {
          const int ib = ch << 4;
          float W[kNfR];
#pragma unroll
          for (int r = 0; r < kNfR; ++r) W[r] = hrow[n0 + r - ib];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float xv = xrow[ib + u];
#pragma unroll
            for (int c = 0; c < kNfR; ++c)
              acc[c] = fmaf(xv, W[(c - u) & 15], acc[c]);
            W[(-u - 1) & 15] = hrow[n0 - (ib + u + 1)];
          }
        }
#pragma unroll
        for (int c = 0; c < kNfR; ++c) yrow[n0 + c] = acc[c];
      }
    }
    __syncthreads();

    // ---- 4. overlap-add + crop + (accumulate) + store ----
    {
      const int t_lo = q0 * frame;
      const int t_hi = min((q0 + p.TFo) * frame, p.N);
      float* outb = p.audio + (size_t)b * p.N;
      for (int t = t_lo + tid; t < t_hi; t += kNfThreads) {
        const int q = t + p.start;                 // index into the OLA buffer
        int j_hi = q / frame;                      // last frame that can reach q
        int j_lo = (q - (p.ylen - 1) + frame - 1) / frame;
        if (q - (p.ylen - 1) < 0) j_lo = 0;
        j_hi = min(j_hi, p.F - 1);
        float acc = 0.f;
        for (int j = j_lo; j <= j_hi; ++j)
          acc += sY[(j - j0) * p.yS + (q - j * frame)];
        if (p.accumulate) acc += outb[t];
        outb[t] = acc;
      }
    }
    __syncthreads();
  }
}

inline int nf_odd(int v) { return v | 1; }

// Fills the derived fields; returns false if the shape is outside the fused path.
inline bool nf_configure(NoiseFusedParams& p, int F, int nb, int N,
                         int window_size) {
  if (nb < 2 || nb > kNfMaxNb) return false;
  p.g = make_ir_geom(nb, window_size);
  p.F = F; p.nb = nb; p.N = N;
  p.frame = (N + F - 1) / F;
  const int S = p.g.S;
  p.start = (S - 1) / 2 - 1;
  if (p.start < 0) return false;
  if (p.frame < 8 || p.frame > 1024) return false;
  p.ylen = p.frame + S - 1;
  p.Hb = (S - 1 - p.start + p.frame - 1) / p.frame;   // ceil((S-1-start)/frame)
  p.Ha = (p.frame - 1 + p.start) / p.frame;
  p.TFo = 32 - p.Hb - p.Ha;
  if (p.TFo < 16) return false;
  p.tiles_per_item = (F + p.TFo - 1) / p.TFo;
  p.mS = nf_odd(nb);
  p.hS = nf_odd(S + 2 * kNfPad);
  p.xS = nf_odd(((p.frame + 15) & ~15) + 16);
  p.yS = nf_odd(((p.ylen + kNfR - 1) / kNfR) * kNfR);
  return nf_smem_layout(p).total <= 200 * 1024;
}

inline bool noise_fused_supported(int F, int nb, int N, int window_size) {
  NoiseFusedParams p;
  return nf_configure(p, F, nb, N, window_size);
}

inline int launch_noise_fused(const float* mags, const float* noise,
                              uint64_t seed, uint64_t offset, float* audio,
                              int B, int F, int nb, int N, int window_size,
                              int accumulate, cudaStream_t st, int raw = 0,
                              float bias = 0.f) {
  NoiseFusedParams p;
  if (!nf_configure(p, F, nb, N, window_size)) {
    set_error("filtered_noise_forward: shape outside the fused path");
    return DDSP_B200_E_UNSUPPORTED;
  }
  p.mags = mags; p.noise = noise; p.audio = audio;
  p.seed = seed; p.offset = offset; p.B = B; p.accumulate = accumulate;
  p.raw = raw; p.bias = bias;
  const long long n_tiles = (long long)B * p.tiles_per_item;
  if (n_tiles >= (1ll << 31)) {
    set_error("filtered_noise_forward: too many tiles");
    return DDSP_B200_E_INVALID;
  }
  p.n_tiles = (int)n_tiles;
  const size_t smem = nf_smem_layout(p).total;
  cudaError_t e = cudaFuncSetAttribute(
      noise_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("filtered_noise_forward: cannot reserve %zu B smem: %s", smem,
              cudaGetErrorString(e));
    return DDSP_B200_E_CUDA;
  }
  const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int grid = (int)std::min<long long>(n_tiles, (long long)kNumSMs * ctas_per_sm);
  noise_fused_kernel<<<grid, kNfThreads, smem, st>>>(p);
  DDSP_CHECK_LAUNCH("filtered_noise_forward(fused)");
  return 0;
}

}  // namespace ddsp
