// Issue-rate probes for the harmonic_v4 inner loop on B200 (sm_100a): FP64 FMA rate,
// and FFMA2 / FADD2 by how many NEW register operands each instruction reads
// (operand reuse between consecutive instructions).  One CTA of 512 threads per SM
// (4 warps per sub-partition); cycles from clock64 on every CTA, averaged.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4000

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// OP 0: DFMA, 8 independent chains
// OP 1: FFMA2 acc[i] = s_i(bcast) * v + acc[i]      (v fixed: slot B reusable) - 16 acc
// OP 2: FFMA2 acc[i] = s(bcast, fixed) * w[i] + acc[i]
// OP 3: FFMA2 acc[i] = s_i(bcast) * w[i & 1] + acc[i]  (v alternates: no reuse) - v4 loop order
// OP 4: FFMA2 acc[i] = p(pair, fixed) * w[i] + acc[i]  (chain form: na pair fixed)
// OP 5: FFMA2 acc[i] = p[i&3](pair) * w[i&7] + acc[i]  (fully packed, nothing reused)
// OP 6: FADD2 acc[i] = acc[i] + w[i & 7]
// OP 7: the v4 group: 8 acc FFMA2 (bcast x) + 4 chain FFMA2 (pair na) + 4 FADD2, source order s0e,t0e,s1e,t1e
// OP 8: the same group, source order s0e,s1e (same v back to back), t0e,t1e
// OP 9: I2F.F64.U32
template <int OP>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float a, float b, int iters) {
  float2 acc[16], w[8];
  float s[16];
  double d[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = f2(a + i, b + i + threadIdx.x * 1e-3f); s[i] = a * (i + 1) * 1e-3f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) { w[i] = f2(a * (i + 1) * 1e-3f, b * (i + 2) * 1e-3f); d[i] = a + i; }
  float2 vo = f2(a, b), ve = f2(b, a), dlo = f2(1e-3f, 2e-3f), dle = f2(3e-3f, 1e-3f), na = f2(-1e-3f * a, -2e-3f * b);
  const double da = a * 1e-3, db = b;
  unsigned u = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (OP == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fma(d[i], da, db);
    }
    if (OP == 9) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { d[i] += (double)(u + i); }
      u += 3;
    }
    if (OP == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(f2(s[i], s[i]), w[0], acc[i]);
    }
    if (OP == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(f2(s[0], s[0]), w[i & 7], acc[i]);
    }
    if (OP == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(f2(s[i], s[i]), w[i & 1], acc[i]);
    }
    if (OP == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(w[0], w[1 + (i % 7)], acc[i]);
    }
    if (OP == 5) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(w[(i + 3) & 7], w[i & 7], acc[i]);
    }
    if (OP == 6) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __fadd2_rn(acc[i], w[i & 7]);
    }
    if (OP == 7 || OP == 8) {
      // x values: s[0..7] as the two float4 rows
      if (OP == 7) {
        acc[0] = __ffma2_rn(f2(s[0], s[0]), vo, acc[0]);
        acc[1] = __ffma2_rn(f2(s[1], s[1]), ve, acc[1]);
        acc[2] = __ffma2_rn(f2(s[4], s[4]), vo, acc[2]);
        acc[3] = __ffma2_rn(f2(s[5], s[5]), ve, acc[3]);
      } else {
        acc[0] = __ffma2_rn(f2(s[0], s[0]), vo, acc[0]);
        acc[2] = __ffma2_rn(f2(s[4], s[4]), vo, acc[2]);
        acc[1] = __ffma2_rn(f2(s[1], s[1]), ve, acc[1]);
        acc[3] = __ffma2_rn(f2(s[5], s[5]), ve, acc[3]);
      }
      dlo = __ffma2_rn(na, vo, dlo); dle = __ffma2_rn(na, ve, dle);
      vo = __fadd2_rn(vo, dlo); ve = __fadd2_rn(ve, dle);
      if (OP == 7) {
        acc[4] = __ffma2_rn(f2(s[2], s[2]), vo, acc[4]);
        acc[5] = __ffma2_rn(f2(s[3], s[3]), ve, acc[5]);
        acc[6] = __ffma2_rn(f2(s[6], s[6]), vo, acc[6]);
        acc[7] = __ffma2_rn(f2(s[7], s[7]), ve, acc[7]);
      } else {
        acc[4] = __ffma2_rn(f2(s[2], s[2]), vo, acc[4]);
        acc[6] = __ffma2_rn(f2(s[6], s[6]), vo, acc[6]);
        acc[5] = __ffma2_rn(f2(s[3], s[3]), ve, acc[5]);
        acc[7] = __ffma2_rn(f2(s[7], s[7]), ve, acc[7]);
      }
      dlo = __ffma2_rn(na, vo, dlo); dle = __ffma2_rn(na, ve, dle);
      vo = __fadd2_rn(vo, dlo); ve = __fadd2_rn(ve, dle);
    }
  }
  long long t1 = clock64();
  float r = vo.x + vo.y + ve.x + ve.y + (float)u;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double inst_per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 512>>>(out, cyc, 1.0001f, 0.9999f, 100);
  cudaDeviceSynchronize();
  k<OP><<<148, 512>>>(out, cyc, 1.0001f, 0.9999f, ITERS);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  // 4 warps per sub-partition each issue inst_per_iter * ITERS instructions
  printf("%-58s %.2f cycles per warp-instruction per sub-partition\n", name,
         avg / (4.0 * inst_per_iter * ITERS));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("DFMA (8 chains)", 16);
  run<9>("I2F.F64.U32 + DADD (+ IADD)", 16);
  run<1>("FFMA2 bcast x_i * v(fixed) + acc_i", 16);
  run<2>("FFMA2 bcast x(fixed) * w_i + acc_i", 16);
  run<3>("FFMA2 bcast x_i * v(alternating) + acc_i", 16);
  run<4>("FFMA2 pair(fixed) * w_i + acc_i", 16);
  run<5>("FFMA2 fully packed, nothing shared", 16);
  run<6>("FADD2 acc_i + w_i", 16);
  run<7>("v4 group (8 acc + 4 chain + 4 FADD2), order o,e,o,e", 16);
  run<8>("v4 group, order o,o,e,e", 16);
  return 0;
}
