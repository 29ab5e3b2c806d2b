// Pipe-throughput microbenchmarks for B200 (sm_100a): ops / clk / SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) bench(float* out, long long* cyc, float a, float b) {
  float x[CHAINS];
  float2 x2[CHAINS];
  uint32_t u[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    x[i] = a + i + threadIdx.x * 1e-3f;
    x2[i] = make_float2(x[i], x[i] + 1.f);
    u[i] = threadIdx.x * 2654435761u + i;
  }
  float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) x[i] = fmaf(x[i], a, b);                       // FFMA 3-reg
      if (OP == 1) x2[i] = __ffma2_rn(x2[i], a2, b2);             // FFMA2
      if (OP == 2) x[i] = __sinf(x[i]);                           // FMUL + MUFU.SIN
      if (OP == 3) u[i] = u[i] * 2654435761u + (uint32_t)it;      // IMAD
      if (OP == 4) x[i] = (float)(int)(__float_as_int(x[i]) + it);// I2F (+IADD)
      if (OP == 5) x[i] = x[i] + a;                               // FADD
      if (OP == 6) x2[i] = __fadd2_rn(x2[i], a2);                 // FADD2
      if (OP == 7) u[i] = (u[i] >> 9) | 0x3F800000u;              // LOP3/SHF
      if (OP == 8) x[i] = fmaf(x[i], 1.0001f, 0.5f);              // FFMA imm
      if (OP == 9) { x[i] = fmaf(x[i], a, b); u[i] = (u[i] + it) ^ 0x55u; } // FFMA + ALU dual
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i] + x2[i].x + x2[i].y + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double ops_per_inst) {
  const int blocks = 148 * 8, threads = 256;
  float* out; long long* cyc;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  bench<OP><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148 * 8]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  // 8 blocks resident per SM, all running concurrently for ~avg cycles
  double lane_ops_per_sm = 8.0 * threads * (double)ITERS * CHAINS * ops_per_inst;
  printf("%-22s %8.3f ms  %10.0f cyc/block  %7.1f lane-ops/clk/SM  (%.1f Tops/s)\n", name, ms,
         avg, lane_ops_per_sm / avg, 148.0 * lane_ops_per_sm / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FFMA (3 reg)", 1);
  run<8>("FFMA (imm)", 1);
  run<1>("FFMA2", 2);
  run<5>("FADD", 1);
  run<6>("FADD2", 2);
  run<2>("__sinf (FMUL+MUFU)", 1);
  run<3>("IMAD", 1);
  run<4>("I2F+IADD", 1);
  run<7>("SHF+LOP3", 1);
  run<9>("FFMA + 2 ALU", 1);
  return 0;
}
