// FFMA2 issue-rate probes for the FIR inner loop (noise_ring.cuh):
//   OP 0: FFMA2 with three packed operands; OP 1: FFMA2 with a broadcast scalar
//   first operand (the R.F32 form the FIR uses); OP 2: scalar FFMA.
// Swept over warps per SM sub-partition (1, 2, 4) with 16 or 33 independent
// accumulators per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int CH>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, float a, float b, int iters) {
  float2 acc[CH];
  float2 w[8];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = make_float2(a + i, b + i + threadIdx.x * 1e-3f);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = make_float2(a * (i + 1), b * (i + 2));
  float s0 = a * 0.5f + threadIdx.x * 1e-6f, s1 = b * 0.25f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (OP == 0) acc[i] = __ffma2_rn(w[(i + 3) & 7], w[i & 7], acc[i]);
      if (OP == 1) acc[i] = __ffma2_rn(make_float2(i & 1 ? s0 : s1, i & 1 ? s0 : s1), w[i & 7], acc[i]);
      if (OP == 2) { acc[i].x = fmaf(s0, w[i & 7].x, acc[i].x); acc[i].y = fmaf(s1, w[i & 7].y, acc[i].y); }
    }
    s0 += 1e-7f;
  }
  long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP, int CH>
void run(const char* name, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  k<OP, CH><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f, 100);
  cudaDeviceSynchronize();
  k<OP, CH><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double inst_per_warp = (double)iters * CH * (OP == 2 ? 2 : 1);
  const int warps_per_smsp = threads / 32 / 4;
  printf("%-28s warps/SMSP %d  acc %2d : %.2f cycles per warp-instruction, %.2f issue cycles per instr per SMSP\n",
         name, warps_per_smsp, CH, c / inst_per_warp, c / inst_per_warp / warps_per_smsp);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int t : {128, 256, 512}) {
    if (t == 128) { run<0, 16>("FFMA2 packed", t); run<1, 16>("FFMA2 scalar-bcast", t); run<2, 16>("FFMA scalar", t);
                    run<0, 33>("FFMA2 packed", t); run<1, 33>("FFMA2 scalar-bcast", t); }
    if (t == 256) { run<0, 16>("FFMA2 packed", t); run<1, 16>("FFMA2 scalar-bcast", t); run<2, 16>("FFMA scalar", t);
                    run<1, 33>("FFMA2 scalar-bcast", t); }
    if (t == 512) { run<0, 16>("FFMA2 packed", t); run<1, 16>("FFMA2 scalar-bcast", t); run<2, 16>("FFMA scalar", t); }
  }
  return 0;
}
